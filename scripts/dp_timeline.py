#!/usr/bin/env python
"""Phase timeline of the fused DP kernels relative to the chain kernel (SSB_CHAIN_TIMELINE=1).
Run under torchrun with 2+ ranks."""
import os
import sys

os.environ["SSB_CHAIN_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from shallowspeed_b200.dataset import synthetic_mnist
from shallowspeed_b200.parallel.comm import ProcessGrid, make_torch_comms
from shallowspeed_b200.parallel.engine import Trainer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
grid = ProcessGrid(world, 1, rank)
dp_comm, pp_comm = make_torch_comms(grid)
SIZES = [784, 128, 127, 126, 125, 124, 123, 10]
x, y = synthetic_mnist(n=128 * world * 4)
xd = torch.from_numpy(x[rank::world].copy()).cuda()
yd = torch.from_numpy(y[rank::world].copy()).cuda()
tr = Trainer(SIZES, global_batch_size=128 * world, dp_comm=dp_comm, pp_comm=pp_comm, grid=grid, use_graph=True)
for i in range(40):
    tr.step_async(xd[(i % 4) * 128:(i % 4 + 1) * 128], yd[(i % 4) * 128:(i % 4 + 1) * 128])
tr.synchronize()
t = tr.engine.chain_timeline()
mma, epi, dp = t[256:512], t[512:768], t[768:1024]
t0 = mma[0]
rel = lambda v: round((v - t0) / 1000.0, 2) if v > 0 else None
if rank == 0:
    print("chain: first mma", rel(mma[0]), " last bwd epilogue seen", rel(max(epi[:64])))
    print("fused DP kernels (CTA 0): start, gemm done, stores issued, fence done, phaseB start, flags seen, phaseC")
    for l in range(7, 0, -1):
        print(f" layer {l}:", [rel(v) for v in dp[8 * l:8 * l + 7]])
dist.barrier()
dist.destroy_process_group()
