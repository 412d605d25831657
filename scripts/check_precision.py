#!/usr/bin/env python
"""Per-parameter error of a few native-engine training steps against the fp32 CPU oracle, for both
tensor-core precisions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shallowspeed_b200.dataset import Dataset, synthetic_mnist
from shallowspeed_b200.layers import MLP
from shallowspeed_b200.optimizer import SGD
from shallowspeed_b200.parallel.engine import NativeWorker
from shallowspeed_b200.pipe import NaiveParallelSchedule, Worker

SIZES = [784, 128, 127, 126, 125, 124, 123, 10]
steps, lr = 3, 0.05
x, y = synthetic_mnist(n=128 * steps)


def run(dev, precision=None):
    model = MLP(SIZES, 0, 1, 128).to(dev)
    opt = SGD(model.parameters(), lr, arena=model.arena)
    ds = Dataset(None, 128, 32, device=dev); ds.local_batch_size = 128; ds.from_arrays(x, y)
    w = Worker(None, None, model, ds, opt) if dev == "cpu" else NativeWorker(None, None, model, ds, opt, precision=precision)
    losses = []
    for b in range(steps):
        w.execute(NaiveParallelSchedule(4, 1, 0), b); losses.append(w.batch_loss())
    if dev != "cpu": w.sync_to_model()
    return [p.data.cpu().double() for p in model.parameters()], losses


init = [p.data.double() for p in MLP(SIZES, 0, 1, 128).parameters()]
ref, lref = run("cpu")
for prec in ("tf32", "fp32"):
    got, l = run("cuda", prec)
    print(prec, "losses", [f"{a:.7f}" for a in l], "cpu", [f"{a:.7f}" for a in lref])
    for i, (p0, a, b) in enumerate(zip(init, got, ref)):
        print(f"  param {i:2d} shape {tuple(a.shape)}: update err {float(((a-p0)-(b-p0)).norm()/((b-p0).norm()+1e-30)):.3e}")
