#!/usr/bin/env python
"""All-reduce through the NVSwitch (multimem.ld_reduce + multimem.st, csrc/kernels/nvls_dp.cu) vs NCCL, same sizes.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/nvls_bench.py

Device-timed (CUDA events, max over ranks), warm-up, prints one JSON line per size on rank 0:
algorithm bandwidth = bytes / time, bus bandwidth = algbw * 2 (n-1) / n (the NCCL convention)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shallowspeed_b200.parallel.comm import ProcessGrid, make_torch_comms  # noqa: E402
from shallowspeed_b200.parallel.engine import make_nvls_context  # noqa: E402
from shallowspeed_b200.utils.timing import max_over_ranks  # noqa: E402


class _Arena:
    """Just enough of ParamArena for make_nvls_context."""

    def __init__(self, numel, dev):
        self.weights = torch.zeros(numel, device=dev)
        self.grads = torch.zeros(numel, device=dev)
        self.numel = numel

    def rebind(self, w, g, copy=True):
        self.weights, self.grads = w, g


class _Model:
    def __init__(self, numel, dev):
        self.arena = _Arena(numel, dev)


def timed(fn, iters, dev):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return max_over_ranks(e0.elapsed_time(e1) / iters, dev)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dp_comm, _ = make_torch_comms(ProcessGrid(world, 1, rank))
    for mb in (1, 4, 16, 64, 256, 1024):
        numel = mb * (1 << 20) // 4
        model = _Model(numel, dev)
        ctx = make_nvls_context(dp_comm, model, 0.0)
        g = model.arena.grads
        g.fill_(float(rank + 1))
        ctx.all_reduce_grads()
        torch.cuda.synchronize(dev)
        ok = bool((g == world * (world + 1) / 2).all())
        ms_nvls = timed(ctx.all_reduce_grads, 20, dev)
        ref = torch.ones(numel, device=dev)
        ms_nccl = timed(lambda: dist.all_reduce(ref), 20, dev)
        if rank == 0:
            nbytes = numel * 4
            bus = 2 * (world - 1) / world
            print(json.dumps({"bytes": nbytes, "n_gpus": world, "nvls_correct": ok,
                              "nvls_ms": round(ms_nvls, 4), "nccl_ms": round(ms_nccl, 4),
                              "nvls_algbw_GBps": round(nbytes / ms_nvls / 1e6, 1), "nccl_algbw_GBps": round(nbytes / ms_nccl / 1e6, 1),
                              "nvls_busbw_GBps": round(bus * nbytes / ms_nvls / 1e6, 1),
                              "nccl_busbw_GBps": round(bus * nbytes / ms_nccl / 1e6, 1)}), flush=True)
        del ctx, model, g
        torch.cuda.synchronize(dev)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
