#!/usr/bin/env python
"""Didactic contrast to the interleaved DP of the engine: NON-overlapped data parallelism
with stock PyTorch autograd - a blocking per-parameter all-reduce AFTER the full backward.

Capability parity with the reference's ``scripts/DDP_PyTorch_MNIST.py`` (:1-167): 3-layer
MLP (hidden 64), rank-strided data subset, CrossEntropy + Adam(1e-3), loss scaled by
1/world, ``all_reduce(param.grad)`` per parameter after ``backward()``, weight-hash sync
assert before and after training, model saved as ``data/models/model_p{world}.pkl`` and
the L1 divergence from the single-process run reported.  torch.distributed (nccl on GPUs,
gloo on CPU) replaces mpi4py.

    python scripts/ddp_blocking_mnist.py                      # 1 process
    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 scripts/ddp_blocking_mnist.py
"""
import argparse
import hashlib
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from shallowspeed_b200.dataset import synthetic_mnist  # noqa: E402


class MLP(nn.Module):
    def __init__(self, hidden=64):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(784, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, 10))

    def forward(self, x):
        return self.net(x)


def model_hash(model):
    h = hashlib.sha1()
    for p in model.parameters():
        h.update(p.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def assert_sync(model, world):
    if world == 1:
        return
    hashes = [None] * world
    dist.all_gather_object(hashes, model_hash(model))
    if len(set(hashes)) > 1:
        raise ValueError("Model hash mismatch")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=128, help="global batch size")
    ap.add_argument("--samples", type=int, default=12800)
    ap.add_argument("--device", default="auto")
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    use_cuda = torch.cuda.is_available() if args.device == "auto" else args.device == "cuda"
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    else:
        torch.set_num_threads(1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29544")
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world)

    torch.manual_seed(0)                                    # identical init on every rank
    model = MLP().to(device)
    assert_sync(model, world)
    x, y = synthetic_mnist(n=args.samples)
    xs = torch.from_numpy(x[rank::world]).to(device)        # rank-strided subset
    ys = torch.from_numpy(y[rank::world]).argmax(1).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = nn.CrossEntropyLoss()
    local_bs = args.batch_size // world
    t0 = time.time()
    for epoch in range(args.epochs):
        for i in range(0, len(xs) - local_bs + 1, local_bs):
            opt.zero_grad()
            loss = loss_fn(model(xs[i:i + local_bs]), ys[i:i + local_bs]) / world
            loss.backward()
            if world > 1:
                for p in model.parameters():                # blocking, one message per parameter
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            opt.step()
        if rank == 0:
            print(f"epoch {epoch}: loss {loss.item() * world:.4f}  time {time.time() - t0:.2f}s")
    assert_sync(model, world)
    if rank == 0:
        out = Path("data/models")
        out.mkdir(parents=True, exist_ok=True)
        torch.save(model.state_dict(), out / f"model_p{world}.pkl")
        ref = out / "model_p1.pkl"
        if world > 1 and ref.exists():
            sd = torch.load(ref, map_location=device)
            div = sum(float((p - sd[k]).abs().sum()) for k, p in model.state_dict().items())
            print(f"L1 divergence from the 1-process model: {div:.6f}")
        print(f"total time {time.time() - t0:.2f}s")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
