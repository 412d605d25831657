"""Opt-in kernels written after the round's GPU budget was spent (split-K GEMMs, lo-twin-refreshing wgrad, multicast
chain kernel).  NOT collected by default (file name): tests/test_gpu_zz_aux.py runs each group in its OWN python process,
so a device trap in one experimental kernel cannot poison the CUDA context of anything else.

    python -m pytest tests/experimental_cases.py -q            # run them directly on a GPU box
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [784, 128, 127, 126, 125, 124, 123, 10]


# ---------------------------------------------------------------------------------------------------------
# Experimental: split-K variant of the FWD / DGRAD GEMMs for wide layers (opt-in: k_splits= / SSB_SPLITK).
# Written after the round's GPU budget was spent, so these are the first executions on hardware;
# each group runs in its own process (tests/test_gpu_zz_aux.py).
# ---------------------------------------------------------------------------------------------------------


def _report(got, ref, tol):
    """One-glance description of a mismatch (these kernels get few GPU runs: make each one count)."""
    err = (got.double() - ref.double()).abs()
    bad = err > tol
    if not bool(bad.any()):
        return "ok"
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    r0, c0 = int(rows[0]), int(cols[0])
    nz = bad & (ref != 0)
    ratio = (got.double()[nz] / ref.double()[nz]).median().item() if bool(nz.any()) else float("nan")
    return (f"max err {err.max().item():.3e} (tol {tol:.1e}); {int(bad.sum())}/{bad.numel()} wrong; rows {int(rows[0])}..{int(rows[-1])} "
            f"({len(rows)} of {got.size(0)}), cols {int(cols[0])}..{int(cols[-1])} ({len(cols)} of {got.size(1)}); "
            f"first bad [{r0},{c0}]: got {got[r0, c0].item():.6g} ref {ref[r0, c0].item():.6g}; median got/ref over bad = {ratio:.4g}; "
            f"zeros in got: {int((got == 0).sum())}, nan: {int(torch.isnan(got).sum())}")


def _rand(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).cuda()


@pytest.mark.parametrize("precision,tol", [("tf32", 3e-3), ("fp32", 2e-5)])
@pytest.mark.parametrize("rows,inp,out,ks", [(32, 2048, 256, -1), (8, 2048, 200, 3), (128, 4096, 128, 8), (33, 1000, 130, 2)])
def test_splitk_forward_matches_oracle(precision, tol, rows, inp, out, ks):
    from shallowspeed_b200.ops import cuda as K

    x, w, b = _rand(rows, inp, seed=1), _rand(out, inp, seed=2) / inp ** 0.5, _rand(out, seed=3)
    ref = torch.relu(x.double() @ w.double().T + b.double()).float()
    plain = K.linear_fwd(x, w, b, relu=True, precision=precision)[:, :out]
    for _ in range(2):      # second launch: the tile counters must have been re-armed
        got = K.linear_fwd(x, w, b, relu=True, precision=precision, k_splits=ks)[:, :out]
        bound = tol * max(ref.abs().max().item(), 1.0) * 4
        assert (got - ref).abs().max().item() <= bound, "vs oracle: " + _report(got, ref, bound)
        assert (got - plain).abs().max().item() <= bound, "vs plain kernel: " + _report(got, plain, bound)


@pytest.mark.parametrize("precision,tol", [("tf32", 3e-3), ("fp32", 2e-5)])
def test_splitk_dgrad_with_relu_mask_matches_oracle(precision, tol):
    from shallowspeed_b200.ops import cuda as K

    rows, inp, out = 32, 300, 4096
    dz, w, act = _rand(rows, out, seed=4), _rand(out, inp, seed=5) / out ** 0.5, _rand(rows, inp, seed=6)
    ref = ((dz.double() @ w.double()) * (act > 0)).float()
    got = K.linear_dgrad(dz, w, mask=act, precision=precision, k_splits=-1)[:, :inp]
    bound = tol * max(ref.abs().max().item(), 1.0) * 4
    assert (got - ref).abs().max().item() <= bound, _report(got, ref, bound)


def test_engine_splitk_optin_trains_like_the_default(monkeypatch):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    sizes = [784, 2048, 2048, 10]
    x, y = synthetic_mnist(n=256)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(sizes, lr=0.05, seed_mode="index")
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(2)]
    monkeypatch.setenv("SSB_SPLITK", "1")
    tr = Trainer(sizes, lr=0.05, seed_mode="index")
    assert "splitk_gemms=0" not in tr.engine.describe()
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(2)]
    assert all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(got, ref))


# ---------------------------------------------------------------------------------------------------------
# Experimental: SGD-fused wgrad that also refreshes the lo twin of the updated weight tile (SSB_FUSE_WLO=1).
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,k,n", [(32, 128, 127), (128, 784, 128), (32, 123, 10), (64, 300, 200)])
def test_wgrad_fused_sgd_refreshes_weight_lo_twin(rows, k, n):
    from shallowspeed_b200.ops import cuda as K

    torch.manual_seed(0)
    lr = 0.05
    dz, x = torch.randn(rows, n, device="cuda"), torch.randn(rows, k, device="cuda")
    ld = (k + 1 + 7) // 8 * 8
    W, G = torch.randn(n, ld, device="cuda"), torch.zeros(n, ld, device="cuda")
    W_ref = W.clone()
    stale = K.lo_twin(W.clone()[:, :k])                    # lo twin of the OLD weights: what a stale read-back would produce
    W_lo = torch.full((n, ld), 7.0, device="cuda")        # sentinel: every weight element must be rewritten
    K.linear_wgrad(dz, x, G[:, :k], accumulate=False, grad_b=G[:, k], weight=W_ref[:, :k], lr=lr, fuse_sgd=True, precision="fp32")
    K.linear_wgrad(dz, x, G[:, :k], accumulate=False, grad_b=G[:, k], weight=W[:, :k], lr=lr, fuse_sgd=True, precision="fp32",
                   weight_lo_out=W_lo[:, :k])
    assert torch.equal(W, W_ref), "update changed: " + _report(W, W_ref, 0.0)
    want = K.lo_twin(W[:, :k])
    assert torch.equal(W_lo[:, :k], want), ("lo twin of the NEW weights, bit for bit: " + _report(W_lo[:, :k], want, 0.0) +
                                            f"; sentinel survivors {int((W_lo[:, :k] == 7.0).sum())}; equals stale twin: {torch.equal(W_lo[:, :k], stale)}")
    assert bool((W_lo[:, k:] == 7.0).all())                # bias slot / padding untouched


def test_engine_fused_weight_lo_optin_is_bitwise_identical(monkeypatch):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    x, y = synthetic_mnist(n=128 * 4)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(SIZES, lr=0.1)
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    monkeypatch.setenv("SSB_FUSE_WLO", "1")
    tr = Trainer(SIZES, lr=0.1)
    assert int(tr.engine.kernels_per_step()) == int(base.engine.kernels_per_step()) - 1   # the split kernel is gone
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    assert got == ref                                      # same products, same order: identical losses
    assert torch.equal(tr.model.arena.weights, base.model.arena.weights)


# ---------------------------------------------------------------------------------------------------------
# Experimental: chain kernel with a 4-CTA cluster sharing the weight stream through TMA multicast (SSB_CHAIN_MC=1).
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("n_mu", [4, 8])
def test_chain_multicast_cluster_is_bitwise_identical(monkeypatch, precision, n_mu):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    x, y = synthetic_mnist(n=128 * 3)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(SIZES, lr=0.1, n_mubatches=n_mu, precision=precision)
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(3)]
    monkeypatch.setenv("SSB_CHAIN_MC", "1")
    tr = Trainer(SIZES, lr=0.1, n_mubatches=n_mu, precision=precision)
    assert tr.engine.uses_chain()
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(3)]
    assert got == ref                                      # same MMAs in the same order: identical losses
    assert torch.equal(tr.model.arena.weights, base.model.arena.weights)



# ---------------------------------------------------------------------------------------------------------
# Experimental: loss values stored straight into pinned host memory (SSB_LOSS_ZEROCOPY=1), no D2H copy node.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env", [{}, {"SSB_NO_CHAIN": "1"}, {"SSB_NO_COALESCE": "1"}])
def test_loss_zero_copy_readback_matches(monkeypatch, env):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    x, y = synthetic_mnist(n=128 * 4)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(SIZES, lr=0.1)
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    monkeypatch.setenv("SSB_LOSS_ZEROCOPY", "1")
    tr = Trainer(SIZES, lr=0.1)
    assert "loss_d2h" not in tr.engine.plan_text(0)
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    assert got == ref
    # pipelined read-back: loss of step i is returned by call i + 1
    tr2 = Trainer(SIZES, lr=0.1)
    lag = [tr2.step_pipelined(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)] + [tr2.flush()]
    assert lag[0] is None and lag[1:] == ref


# ---------------------------------------------------------------------------------------------------------
# Experimental: all layers' weight-gradient tiles in ONE launch (SSB_WGRAD_GROUP=1); with the lo-twin-refreshing
# epilogue and the zero-copy loss the whole step is two graph nodes.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_wgrad_group_launch_is_bitwise_identical(monkeypatch, precision):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer

    x, y = synthetic_mnist(n=128 * 4)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(SIZES, lr=0.1, precision=precision)
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    monkeypatch.setenv("SSB_WGRAD_GROUP", "1")
    tr = Trainer(SIZES, lr=0.1, precision=precision)
    plan = tr.engine.plan_text(0)
    assert "wgrad_group" in plan and " gemm " not in plan
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    assert got == ref
    assert torch.equal(tr.model.arena.weights, base.model.arena.weights)


def test_two_node_step_all_optins_together(monkeypatch):
    from shallowspeed_b200.dataset import synthetic_mnist
    from shallowspeed_b200.parallel.engine import Trainer
    from shallowspeed_b200.parallel.plan_check import check_plan

    x, y = synthetic_mnist(n=128 * 4)
    xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    base = Trainer(SIZES, lr=0.1)
    ref = [base.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    for k in ("SSB_WGRAD_GROUP", "SSB_FUSE_WLO", "SSB_LOSS_ZEROCOPY"):
        monkeypatch.setenv(k, "1")
    tr = Trainer(SIZES, lr=0.1)
    stats = check_plan(tr.engine.plan_text(0))
    assert stats["kernels_and_copies"] == 2, tr.engine.plan_text(0)      # chain kernel + grouped wgrad, nothing else
    got = [tr.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    assert got == ref and torch.equal(tr.model.arena.weights, base.model.arena.weights)
    # ... and with the multicast chain kernel on top (only meaningful once test_chain_multicast_* passes on its own)
    monkeypatch.setenv("SSB_CHAIN_MC", "1")
    tr2 = Trainer(SIZES, lr=0.1)
    got2 = [tr2.step(xh[i * 128:(i + 1) * 128], yh[i * 128:(i + 1) * 128]) for i in range(4)]
    assert got2 == ref, "two-node step is fine, the multicast chain kernel on top of it is not"
