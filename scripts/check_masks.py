import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shallowspeed_b200.dataset import synthetic_mnist
from shallowspeed_b200.layers import MLP
SIZES = [784, 128, 127, 126, 125, 124, 123, 10]
x, y = synthetic_mnist(n=128)
xc = torch.from_numpy(x)
cpu = MLP(SIZES, 0, 1, 128); gpu = MLP(SIZES, 0, 1, 128).to("cuda")
a_c, a_g = xc, xc.cuda()
for i, (lc, lg) in enumerate(zip(cpu.linears, gpu.linears)):
    a_c = lc.forward(a_c, 0); a_g = lg.forward(a_g, 0)
    ag = a_g.cpu()
    mism = ((a_c > 0) != (ag > 0))
    err = float((a_c - ag).abs().max() / a_c.abs().max())
    vals = torch.maximum(a_c.abs(), ag.abs())[mism]
    print(f"layer {i+1}: max rel err {err:.2e}  mask mismatches {int(mism.sum())} of {mism.numel()}  |y| at mismatches {[f'{v:.1e}' for v in vals[:5].tolist()]}  exact zeros pre-act? n(y==0 both)={int(((a_c==0)&(ag==0)).sum())}")
