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
LR = 0.05
x, y = synthetic_mnist(n=128)
m = MLP(SIZES, 0, 1, 128)
W = [l._params["W"].data.double() for l in m.linears]; B = [l._params["b"].data.double() for l in m.linears]
dW = [torch.zeros_like(w) for w in W]; dB = [torch.zeros_like(b) for b in B]
for mu in range(4):
    a = [torch.from_numpy(x[mu * 32:(mu + 1) * 32]).double()]
    for i in range(7):
        z = a[-1] @ W[i].T + B[i]
        a.append(z.clamp_min(0) if i < 6 else z)
    p = F.softmax_ref(a[7]); t = torch.from_numpy(y[mu * 32:(mu + 1) * 32]).double()
    dz = F.softmax_grad_ref(F.mse_loss_grad_ref(p, t, 128), a[7])
    for l in range(7, 0, -1):
        dW[l - 1] += dz.T @ a[l - 1]; dB[l - 1] += dz.sum(0, keepdim=True)
        if l > 1:
            dz = (dz @ W[l - 1]) * (a[l - 1] > 0)
gm = MLP(SIZES, 0, 1, 128).to("cuda")
ds = Dataset(None, 128, 32, device="cuda"); ds.local_batch_size = 128; ds.from_arrays(x, y)
w = NativeWorker(None, None, gm, ds, SGD(gm.parameters(), LR, arena=gm.arena), precision=os.environ.get("PREC", "fp32"))
w.execute(NaiveParallelSchedule(4, 1, 0), 0); w.sync_to_model()
for i, lin in enumerate(gm.linears):
    Wg = lin._params["W"].data.cpu().double(); Bg = lin._params["b"].data.cpu().double()
    uw = (Wg - W[i]) / -LR; ub = (Bg - B[i]) / -LR
    print(f"layer {i+1}: dW err {float((uw - dW[i]).norm() / dW[i].norm()):.2e}   db err {float((ub - dB[i]).norm() / dB[i].norm()):.2e}   |dW| {float(dW[i].norm()):.3e} |W| {float(W[i].norm()):.2e}")
