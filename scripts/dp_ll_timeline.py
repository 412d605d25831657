#!/usr/bin/env python
"""Phase timeline of the LL two-shot data-parallel kernel relative to the chain kernel (SSB_CHAIN_TIMELINE=1).
Run under torchrun with 2, 4 or 8 ranks; rank 0 prints, one line per tile (the table order is layer L .. 1)."""
import json
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
mma, epi, ll = t[256:512], t[512:768], t[768:1024]
t0 = mma[0]
rel = lambda v: round((v - t0) / 1000.0, 2) if v > 0 else None
if rank == 0:
    chain_end = max(epi[:64])
    print(json.dumps({"dp": world, "chain_first_mma_us": rel(mma[0]), "chain_last_dgrad_epilogue_us": rel(chain_end)}))
    # tile order in the table: layers 7..2 have 4 tiles each (in / 32), layer 1 has 25; only the first 28 tiles are stamped
    names = [f"L{7 - k // 4} tile {k % 4}" for k in range(24)] + [f"L1 tile {k}" for k in range(4)]
    print("tile: resident | operands landed | accumulator done | phase A issued | phase B done | phase C done   (us, chain kernel's first MMA = 0)")
    for k, name in enumerate(names):
        st = ll[8 * k:8 * k + 6]
        if st[0]:
            print(f" {name:10s}:", [rel(v) for v in st], "  after chain end:", round((st[5] - chain_end) / 1000.0, 2))
dist.barrier()
dist.destroy_process_group()
