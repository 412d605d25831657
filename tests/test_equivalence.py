"""End-to-end equivalence: sequential == DP == PP == DP x PP for every schedule (what the
reference can only eyeball through accuracy curves, SURVEY.md section 4 "gaps").  Runs the
real Worker VM on an in-process thread fabric (fast) and once over torch.distributed/gloo
with spawned processes (the real multi-process path)."""
import os
import threading

import pytest
import torch

from shallowspeed_b200.dataset import Dataset, synthetic_mnist
from shallowspeed_b200.layers import MLP
from shallowspeed_b200.optimizer import SGD
from shallowspeed_b200.parallel.comm import ProcessGrid, ThreadFabric
from shallowspeed_b200.pipe import (GPipeSchedule, InferenceSchedule, NaiveParallelSchedule,
                                    PipeDreamSchedule, Worker)
from shallowspeed_b200.utils import assert_sync, get_model_hash

SIZES = [784, 128, 127, 126, 125, 124, 123, 10]
GBS, N_MU, LR, STEPS, N_SAMPLES = 128, 4, 0.05, 6, 128 * 6
_X, _Y = synthetic_mnist(n=N_SAMPLES)


def _make_dataset(dp_rank, dp, mubatch):
    ds = Dataset(None, GBS, mubatch)
    ds.local_batch_size = GBS // dp
    return ds.from_arrays(_X[dp_rank::dp], _Y[dp_rank::dp])


def run_layout(dp, pp, sched_cls, n_mu=N_MU, steps=STEPS):
    """returns {stage: [param tensors]} of replica 0 and the set of per-stage hashes."""
    world = dp * pp
    dp_fabrics = [ThreadFabric(dp) for _ in range(pp)]
    pp_fabrics = [ThreadFabric(pp) for _ in range(dp)]
    results, errors = {}, []

    def rank_main(rank):
        try:
            torch.set_num_threads(1)
            grid = ProcessGrid(dp, pp, rank)
            dp_comm = dp_fabrics[grid.stage].comm(grid.replica)
            pp_comm = pp_fabrics[grid.replica].comm(grid.stage)
            model = MLP(SIZES, grid.stage, pp, GBS)
            opt = SGD(model.parameters(), LR, arena=model.arena)
            ds = _make_dataset(grid.replica, dp, GBS // dp // n_mu)
            worker = Worker(dp_comm, pp_comm, model, ds, opt)
            sched = sched_cls(n_mu, pp, grid.stage)
            for b in range(steps):
                worker.execute(sched, b)
            assert_sync(dp_comm, get_model_hash(model))
            results[rank] = [p.data.clone() for p in model.parameters()]
        except Exception as e:  # pragma: no cover
            errors.append((rank, e))
            for f in dp_fabrics + pp_fabrics:
                f.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join(120) for t in threads]
    assert not errors, errors
    # replicas must be bit-identical
    for rank in range(world):
        g = ProcessGrid(dp, pp, rank)
        for a, b in zip(results[rank], results[g.stage]):
            assert torch.equal(a, b), "DP replicas diverged"
    return [p for s in range(pp) for p in results[s]]


@pytest.fixture(scope="module")
def sequential():
    return run_layout(1, 1, NaiveParallelSchedule)


def _close(a, b, tol=2e-6):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert float((x - y).abs().max()) < tol, float((x - y).abs().max())


@pytest.mark.parametrize("dp,pp,cls", [
    (1, 1, GPipeSchedule), (1, 1, PipeDreamSchedule),
    (2, 1, NaiveParallelSchedule), (4, 1, GPipeSchedule), (8, 1, PipeDreamSchedule),
    (1, 2, NaiveParallelSchedule), (1, 2, GPipeSchedule), (1, 2, PipeDreamSchedule),
    (1, 4, NaiveParallelSchedule), (1, 4, GPipeSchedule), (1, 4, PipeDreamSchedule),
    (2, 2, GPipeSchedule), (2, 4, GPipeSchedule), (4, 2, PipeDreamSchedule), (2, 2, NaiveParallelSchedule),
])
def test_layout_matches_sequential(sequential, dp, pp, cls):
    _close(run_layout(dp, pp, cls), sequential)


@pytest.mark.parametrize("n_mu", [1, 2, 8])
def test_microbatch_count_does_not_change_the_update(sequential, n_mu):
    _close(run_layout(1, 1, GPipeSchedule, n_mu=n_mu), sequential)
    _close(run_layout(1, 2, PipeDreamSchedule, n_mu=n_mu), sequential)


def test_runs_are_bit_deterministic():
    a, b = run_layout(2, 2, PipeDreamSchedule, steps=3), run_layout(2, 2, PipeDreamSchedule, steps=3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_training_reduces_loss_and_inference_schedule_works():
    model = MLP(SIZES, 0, 1, GBS)
    opt = SGD(model.parameters(), 0.5, arena=model.arena)
    ds = _make_dataset(0, 1, GBS // N_MU)
    worker = Worker(None, None, model, ds, opt)
    losses = []
    for epoch in range(6):
        for b in range(ds.get_num_batches()):
            worker.execute(NaiveParallelSchedule(N_MU, 1, 0), b)
            losses.append(worker.batch_loss())
    assert losses[-1] < 0.8 * losses[0]
    val = Dataset(None, GBS, GBS)
    val.from_arrays(_X[:256], _Y[:256])
    vw = Worker(None, None, model, val, None)
    model.eval()
    vw.execute(InferenceSchedule(1, 1, 0), 0)
    probs = vw.output_buffers[0]
    assert probs.shape == (GBS, 10) and torch.allclose(probs.sum(1), torch.ones(GBS), atol=1e-4)
    model.train()


# ------------------------------------------------------------------ real processes over gloo
def _gloo_main(rank, dp, pp, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(dp * pp))
    import torch.distributed as dist

    from shallowspeed_b200.parallel.comm import make_torch_comms

    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=dp * pp)
    grid = ProcessGrid(dp, pp, rank)
    dp_comm, pp_comm = make_torch_comms(grid)
    model = MLP(SIZES, grid.stage, pp, GBS)
    opt = SGD(model.parameters(), LR, arena=model.arena)
    ds = _make_dataset(grid.replica, dp, GBS // dp // N_MU)
    worker = Worker(dp_comm, pp_comm, model, ds, opt)
    sched = PipeDreamSchedule(N_MU, pp, grid.stage)
    for b in range(STEPS):
        worker.execute(sched, b)
    assert_sync(dp_comm, get_model_hash(model))
    if grid.replica == 0:
        torch.save([p.data.clone() for p in model.parameters()], os.path.join(out_dir, f"stage{grid.stage}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_multiprocess_dp2_pp2_matches_sequential(sequential, tmp_path):
    import torch.multiprocessing as mp

    port = 29600 + os.getpid() % 300
    mp.spawn(_gloo_main, args=(2, 2, port, str(tmp_path)), nprocs=4, join=True)
    got = [p for s in range(2) for p in torch.load(tmp_path / f"stage{s}.pt")]
    _close(got, sequential)
