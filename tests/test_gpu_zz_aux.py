"""Auxiliary subsystems on a real GPU: step watchdog, exposed-communication accounting (runs last)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [784, 128, 127, 126, 125, 124, 123, 10]


def test_trainer_watchdog_passes_through_when_healthy():
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    x, y = synthetic_mnist(n=256)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    plain, guarded = Trainer(SIZES, lr=0.1), Trainer(SIZES, lr=0.1, watchdog_s=60.0)
    for i in range(2):
        a = plain.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128])
        b = guarded.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128])
        assert a == b
    assert guarded.engine.wait(1.0) and guarded.engine.comm_status() == ""


def test_comm_timing_mode_is_eager_and_reports_zero_without_communication(monkeypatch):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    x, y = synthetic_mnist(n=128)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    ref = Trainer(SIZES, lr=0.1).step(xh, yh)
    monkeypatch.setenv("SSB_COMM_TIMING", "1")
    tr = Trainer(SIZES, lr=0.1)
    assert tr.engine.comm_timing_enabled() and int(tr.engine.graph_nodes()) == 0
    assert tr.step(xh, yh) == ref
    assert tr.engine.comm_timing() == (0.0, 0.0)



@pytest.mark.parametrize("env", [{}, {"SSB_NO_COALESCE": "1"}, {"SSB_NO_CHAIN": "1"}, {"SSB_NO_CHAIN": "1", "SSB_NO_COALESCE": "1"}])
def test_lowered_plans_pass_the_self_check(monkeypatch, env):
    """Every wait refers to an earlier record, no event is recorded twice, every side stream is forked and joined."""
    from shallowspeed_b200.parallel.engine import Trainer
    from shallowspeed_b200.parallel.plan_check import check_plan

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    tr = Trainer(SIZES, lr=0.1)
    for s in (0, 1):
        stats = check_plan(tr.engine.plan_text(s))
        assert stats["kernels_and_copies"] >= 2 and stats["ops"] == len(tr.engine.plan_text(s).splitlines())
