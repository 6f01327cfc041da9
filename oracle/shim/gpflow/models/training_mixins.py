"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.models.training_mixins [3P-memory]; SURVEY.md A.8."""
from typing import Any

Data = Any


class ExternalDataTrainingLossMixin:
    def training_loss(self, data):
        return self._training_loss(data)

    def training_loss_closure(self, data, *, compile=True):
        training_loss = self.training_loss
        if hasattr(data, "__next__"):
            def closure():
                return training_loss(next(data))
        else:
            def closure():
                return training_loss(data)
        return closure
