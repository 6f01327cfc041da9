"""Data-parallel plumbing over the points: shard layout and the single all-reduce.

Every rank holds a contiguous shard of (X, Y) and replicas of all parameters; the per-shard sums produced by
`mgp_elbo_local` are combined with ONE all-reduce(sum) of the flat reduce buffer, after which every rank runs the
replicated `mgp_elbo_finish`.  Two ways to run the collective:

  * "torch": torch.distributed's all_reduce between the two C calls (NCCL over NVLink on the GPU box; gloo in the CPU
    tests of this module's logic);
  * "nccl": an ncclComm_t created here with the NCCL torch already has loaded and handed to the context
    (`mgp_ctx_set_comm`): libmgp then issues ncclAllReduce itself, on its own stream, per layer as soon as that layer's
    partial sums are folded — no torch.distributed call on the step's path (SURVEY.md §8b's `nccl_comm_or_null`).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`: the first n_total % world ranks get one extra row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_count_and_offset(n_local: int, group=None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """(global number of points, global index of this rank's first point) from one small all-reduce."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = int(n_local)
    dist.all_reduce(counts, group=group)
    return int(counts.sum()), int(counts[:rank].sum())


def all_reduce_sum_(buf: torch.Tensor, group=None) -> torch.Tensor:
    """The step's single fused collective: in-place sum of the flat float64 reduce buffer across ranks."""
    import torch.distributed as dist
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


# ---- an ncclComm_t of our own, for the C-ABI path ---------------------------------------------------------------
class _NcclUniqueId(__import__("ctypes").Structure):
    _fields_ = [("internal", __import__("ctypes").c_char * 128)]


_nccl = None


def _nccl_lib():
    """The libnccl.so.2 torch's CUDA build has already loaded (same soname -> same instance; libmgp resolves
    ncclAllReduce from it too)."""
    global _nccl
    if _nccl is None:
        import ctypes as C
        lib = C.CDLL("libnccl.so.2")
        lib.ncclGetUniqueId.restype, lib.ncclGetUniqueId.argtypes = C.c_int, [C.POINTER(_NcclUniqueId)]
        lib.ncclCommInitRank.restype = C.c_int
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        lib.ncclCommDestroy.restype, lib.ncclCommDestroy.argtypes = C.c_int, [C.c_void_p]
        lib.ncclGetErrorString.restype, lib.ncclGetErrorString.argtypes = C.c_char_p, [C.c_int]
        _nccl = lib
    return _nccl


def create_nccl_comm(group=None, device: Optional[torch.device] = None) -> int:
    """ncclComm_t (as an int) spanning the ranks of `group`: rank 0's ncclUniqueId travels through torch.distributed
    (any backend), every rank then calls ncclCommInitRank on its current CUDA device."""
    import ctypes as C
    import torch.distributed as dist
    lib = _nccl_lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = _NcclUniqueId()
    if rank == 0:
        rc = lib.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError("ncclGetUniqueId: " + lib.ncclGetErrorString(rc).decode())
    backend = dist.get_backend(group)
    dev = device if (backend == "nccl" and device is not None) else torch.device("cpu")
    # (a c_char array FIELD reads back as bytes truncated at the first NUL: take the raw 128 bytes of the struct)
    raw = C.string_at(C.addressof(uid), 128) if rank == 0 else bytes(128)
    box = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    C.memmove(C.addressof(uid), box.cpu().numpy().tobytes(), 128)
    comm = C.c_void_p()
    rc = lib.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0 or not comm.value:
        raise RuntimeError("ncclCommInitRank: " + lib.ncclGetErrorString(rc).decode())
    return int(comm.value)


def destroy_nccl_comm(comm: int) -> None:
    if comm:
        import ctypes as C
        _nccl_lib().ncclCommDestroy(C.c_void_p(comm))
