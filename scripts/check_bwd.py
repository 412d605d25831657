import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shallowspeed_b200.dataset import Dataset, synthetic_mnist
from shallowspeed_b200.layers import MLP
from shallowspeed_b200.optimizer import SGD
from shallowspeed_b200.ops import functional as F
from shallowspeed_b200.parallel.engine import NativeWorker
from shallowspeed_b200.pipe import NaiveParallelSchedule
SIZES = [784, 128, 127, 126, 125, 124, 123, 10]
x, y = synthetic_mnist(n=128)
# ---- CPU, fp64 truth for micro-batch 0 (rows 0..31)
m = MLP(SIZES, 0, 1, 128)
W = [l._params["W"].data.double() for l in m.linears]; B = [l._params["b"].data.double() for l in m.linears]
a = [torch.from_numpy(x[:32]).double()]
for i in range(7):
    z = a[-1] @ W[i].T + B[i]
    a.append(z.clamp_min(0) if i < 6 else z)
p = F.softmax_ref(a[7]); t = torch.from_numpy(y[:32]).double()
dz = [None] * 8
dz[7] = F.softmax_grad_ref(F.mse_loss_grad_ref(p, t, 128), a[7])
for l in range(7, 1, -1):
    dz[l - 1] = (dz[l] @ W[l - 1]) * (a[l - 1] > 0)
# ---- engine
gm = MLP(SIZES, 0, 1, 128).to("cuda")
ds = Dataset(None, 128, 32, device="cuda"); ds.local_batch_size = 128; ds.from_arrays(x, y)
w = NativeWorker(None, None, gm, ds, SGD(gm.parameters(), 0.0, arena=gm.arena), precision=os.environ.get("PREC", "fp32"))
sched = NaiveParallelSchedule(4, 1, 0)
w.execute(sched, 0); w.synchronize()
eng = w.engine_for(sched)
for l in range(1, 8):
    ea = float((eng.act(0, l).cpu().double() - a[l]).norm() / a[l].norm())
    ed = float((eng.dz(0, l).cpu().double() - dz[l]).norm() / dz[l].norm())
    print(f"layer {l}: act err {ea:.2e}   dz err {ed:.2e}")
