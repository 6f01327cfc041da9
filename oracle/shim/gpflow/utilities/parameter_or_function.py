"""TEST INFRASTRUCTURE ONLY (oracle shim)."""
from typing import Any

ParameterOrFunction = Any
