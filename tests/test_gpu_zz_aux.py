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
        assert stats["kernels_and_copies"] >= 3 and stats["ops"] == len(tr.engine.plan_text(s).splitlines())


# ---------------------------------------------------------------------------------------------------------
# Experimental opt-in kernels (tests/experimental_cases.py): one isolated python process per group, so a device
# trap in a kernel that has never run on hardware cannot poison this process.  Non-strict xfail: the outcome is
# information for the next round, not a gate.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.xfail(strict=False, reason="opt-in kernels written without GPU access: first run on hardware")
@pytest.mark.parametrize("group", ["splitk", "weight_lo", "chain_multicast", "loss_zero_copy", "wgrad_group_launch or two_node_step"])
def test_experimental_group_in_isolated_process(group):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "experimental_cases.py"), "-k", group, "-q", "-x",
                        "-p", "no:cacheprovider"], cwd=root, capture_output=True, text=True, timeout=300)
    tail = (r.stdout or "")[-3000:] + (r.stderr or "")[-1500:]
    assert r.returncode == 0, tail
