"""Numerics of the functional ops on the CPU (oracle) path: shape contracts and central
finite-difference gradient checks in fp64 - the strategy of the reference's
tests/test_functional.py, written against torch tensors."""
import pytest
import torch

from shallowspeed_b200 import functional as F

EPS = 1e-5
torch.manual_seed(0)


def numeric_grad(fn, x, upstream):
    """d/dx sum(fn(x) * upstream) by central differences (fp64)."""
    g = torch.zeros_like(x)
    flat, gflat = x.view(-1), g.view(-1)
    for i in range(flat.numel()):
        old = flat[i].item()
        flat[i] = old + EPS
        hi = (fn(x) * upstream).sum().item()
        flat[i] = old - EPS
        lo = (fn(x) * upstream).sum().item()
        flat[i] = old
        gflat[i] = (hi - lo) / (2 * EPS)
    return g


def test_shapes():
    x, w, b = torch.randn(5, 7), torch.randn(3, 7), torch.randn(1, 3)
    y = F.linear(x, w, b)
    assert y.shape == (5, 3)
    dx, dw, db = F.linear_grad(torch.randn(5, 3), x, w)
    assert dx.shape == x.shape and dw.shape == w.shape and db.shape == (3,)
    assert F.relu(x).shape == x.shape
    assert F.relu_grad(x, x > 0).shape == x.shape
    assert F.softmax(x).shape == x.shape
    assert F.softmax_grad(x, x).shape == x.shape
    assert F.mse_loss_grad(x, x, 5).shape == x.shape


def test_relu_values():
    x = torch.tensor([[-1.0, 0.0, 2.0]])
    assert torch.equal(F.relu(x), torch.tensor([[0.0, 0.0, 2.0]]))
    assert torch.equal(F.relu_grad(torch.ones_like(x), x > 0), torch.tensor([[0.0, 0.0, 1.0]]))


def test_relu_grad_fd():
    x = torch.randn(4, 6, dtype=torch.float64)
    x[x.abs() < 1e-3] = 0.5
    up = torch.randn(4, 6, dtype=torch.float64)
    assert torch.allclose(F.relu_grad(up, x > 0), numeric_grad(F.relu, x.clone(), up), atol=1e-7)


def test_linear_grad_fd():
    x = torch.randn(5, 7, dtype=torch.float64)
    w = torch.randn(3, 7, dtype=torch.float64)
    b = torch.randn(1, 3, dtype=torch.float64)
    up = torch.randn(5, 3, dtype=torch.float64)
    dx, dw, db = F.linear_grad(up, x, w)
    assert torch.allclose(dx, numeric_grad(lambda t: F.linear(t, w, b), x.clone(), up), atol=1e-6)
    assert torch.allclose(dw, numeric_grad(lambda t: F.linear(x, t, b), w.clone(), up), atol=1e-6)
    assert torch.allclose(db.reshape(1, -1), numeric_grad(lambda t: F.linear(x, w, t), b.clone(), up), atol=1e-6)


def test_softmax_properties():
    x = torch.randn(6, 10)
    y = F.softmax(x)
    assert (y > 0).all()
    assert torch.allclose(y.sum(dim=1), torch.ones(6), atol=1e-5)
    assert torch.allclose(F.softmax(x + 3.0), y, atol=1e-6)  # shift invariance


def test_softmax_global_max_and_eps():
    # contract kept from the reference (functional.py:24-27): shift by the GLOBAL max of
    # the micro-batch and add 1e-7 to the denominator
    x = torch.tensor([[0.0, 1.0], [10.0, 11.0]], dtype=torch.float64)
    e = torch.exp(x - 11.0)
    expect = e / (e.sum(dim=1, keepdim=True) + 1e-7)
    assert torch.allclose(F.softmax(x), expect, atol=0, rtol=1e-12)


@pytest.mark.parametrize("rows", [1, 4])
def test_softmax_grad_fd(rows):
    # with several rows the global max couples rows only through the constant shift,
    # whose derivative vanishes up to the 1e-7 epsilon -> same Jacobian row by row
    x = torch.randn(rows, 10, dtype=torch.float64)
    up = torch.randn(rows, 10, dtype=torch.float64)
    assert torch.allclose(F.softmax_grad(up, x), numeric_grad(F.softmax, x.clone(), up), atol=1e-6)


def test_mse_loss_and_grad_fd():
    x = torch.randn(4, 10, dtype=torch.float64)
    t = torch.randn(4, 10, dtype=torch.float64)
    assert F.mse_loss(t, t, 4).item() == 0.0
    assert abs(F.mse_loss(x, t, 8).item() - ((t - x) ** 2).sum().item() / 8) < 1e-12
    g = numeric_grad(lambda z: F.mse_loss(z, t, 8).reshape(1, 1), x.clone(), torch.ones(1, 1, dtype=torch.float64))
    assert torch.allclose(F.mse_loss_grad(x, t, 8), g, atol=1e-6)


def test_loss_head_fused_matches_chain():
    z = torch.randn(8, 10, dtype=torch.float64)
    t = torch.zeros(8, 10, dtype=torch.float64)
    t[torch.arange(8), torch.randint(0, 10, (8,))] = 1
    dz, p, loss = F.loss_head_backward(z, t, 32)
    assert torch.allclose(p, F.softmax(z))
    assert torch.allclose(dz, F.softmax_grad(F.mse_loss_grad(p, t, 32), z), atol=1e-12)
    assert abs(float(loss) - float(F.mse_loss(p, t, 32))) < 1e-12
