"""Utilities.  Parity: ``rprint``, ``get_model_hash``, ``assert_sync``
(reference ``shallowspeed/utils.py:8-31``) + timing/logging/clock helpers the reference
lacks (tracing, metrics)."""
from __future__ import annotations

import os
from hashlib import sha1


def _world_rank() -> int:
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0"))


def rprint(*args, **kwargs):
    """print on world rank 0 only."""
    if _world_rank() == 0:
        print(*args, **kwargs)


def get_model_hash(model) -> str:
    """SHA-1 over the concatenated per-parameter SHA-1 digests (content hash of the
    weights; device tensors are copied to the host - this is a debug check, not a hot
    op)."""
    hash_str = ""
    for param in model.parameters():
        data = param.data if hasattr(param, "data") and hasattr(param, "grad") else param
        arr = data.detach().to("cpu").contiguous().numpy()
        hash_str += sha1(arr.tobytes()).hexdigest()
    return sha1(hash_str.encode("utf-8")).hexdigest()


def assert_sync(comm, model_hash):
    """All DP replicas must hold bit-identical weights (reference train.py:155)."""
    hashes = comm.gather(model_hash, root=0)
    if comm.Get_rank() == 0 and len(set(hashes)) > 1:
        raise ValueError("Model hash mismatch")


from .timing import CudaTimer, ClockSampler, StepLogger  # noqa: E402

__all__ = ["rprint", "get_model_hash", "assert_sync", "CudaTimer", "ClockSampler", "StepLogger"]
