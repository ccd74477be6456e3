"""MK-MMD / Deep-MMD losses: vectorised kernels vs naive per-pair loops, in-house QP vs SciPy, beta optimisation
behaviour (scenarios mirror the reference's tests/losses/test_mkmmd_loss.py:465-620, values re-derived here)."""

import math

import numpy as np
import pytest
import torch
from scipy import optimize

from fl4health_b200.losses.deep_mmd_loss import DeepMmdLoss
from fl4health_b200.losses.mkmmd_loss import MkMmdLoss, solve_simplex_like_qp

DEVICE = torch.device("cpu")
X = torch.tensor([[1, 1, 1], [3, 4, 4], [4, 2, 1], [2, 1, 4], [1, 2, 1], [3, 4, 4], [4, 3, 3], [3, 3, 2], [4, 4, 4], [4, 2, 1], [1, 1, 1]],
                 dtype=torch.float32)
Y = torch.tensor([[4, 3, 4], [1, 2, 2], [3, 4, 1], [1, 4, 2], [4, 2, 4], [4, 1, 2], [2, 2, 1], [2, 3, 4], [3, 2, 1], [4, 1, 4], [2, 2, 2]],
                 dtype=torch.float32)
GAMMAS = torch.tensor([2.0, 1.0, 0.5])


def _rbf(a: torch.Tensor, b: torch.Tensor, gamma: float) -> float:
    return math.exp(-float(((a - b) ** 2).sum()) / gamma)


def test_full_estimator_matches_naive_loops() -> None:
    loss = MkMmdLoss(DEVICE, gammas=GAMMAS)
    h = loss.compute_all_h_u_all_samples(X, Y)
    n = X.shape[0]
    for k, gamma in enumerate(GAMMAS.tolist()):
        for s in (0, 3, 10):
            for t in (1, 3, 7):
                naive = _rbf(X[s], X[t], gamma) + _rbf(Y[s], Y[t], gamma) - _rbf(X[s], Y[t], gamma) - _rbf(Y[s], X[t], gamma)
                assert h[k, s, t].item() == pytest.approx(naive, abs=1e-5)
    hat_d = loss.compute_hat_d_per_kernel(h)
    assert hat_d.shape == (3, 1) and hat_d[1, 0].item() == pytest.approx(h[1].mean().item())
    betas = torch.tensor([1.5, 2.0, -1.0]).reshape(-1, 1)
    assert loss.compute_mkmmd(X, Y, betas).item() == pytest.approx(float((betas * hat_d).sum()), abs=1e-5)
    q = loss.compute_hat_q_k(h, hat_d)
    centered = h - hat_d.reshape(3, 1, 1)
    assert q[0, 2].item() == pytest.approx(float((centered[0] * centered[2]).sum()) / (n * n - 1), abs=1e-6)
    assert torch.allclose(q, q.t())


def test_linear_estimator_matches_naive_loops() -> None:
    loss = MkMmdLoss(DEVICE, gammas=GAMMAS, perform_linear_approximation=True)
    quads = loss.construct_quadruples(X, Y)
    assert quads.shape == (5, 4, 3) and torch.equal(quads[1, 0], X[2]) and torch.equal(quads[1, 3], Y[3])
    h = loss.compute_all_h_u_linear(X, Y)
    assert h.shape == (3, 5)
    for k, gamma in enumerate(GAMMAS.tolist()):
        for i in range(5):
            x1, x2, y1, y2 = X[2 * i], X[2 * i + 1], Y[2 * i], Y[2 * i + 1]
            naive = _rbf(x1, x2, gamma) + _rbf(y1, y2, gamma) - _rbf(x1, y2, gamma) - _rbf(x2, y1, gamma)
            assert h[k, i].item() == pytest.approx(naive, abs=1e-5)
    delta = loss.form_h_u_delta_w_i(h)
    assert delta.shape == (3, 2) and delta[0, 1].item() == pytest.approx((h[0, 2] - h[0, 3]).item())
    q = loss.compute_hat_q_k_linear(h)
    assert q[1, 2].item() == pytest.approx(float((delta[1] * delta[2]).sum()) / 2, abs=1e-6)


def test_defaults_and_normalisation() -> None:
    loss = MkMmdLoss(DEVICE, normalize_features=True)
    assert loss.kernel_num == 19 and loss.gammas[0].item() == pytest.approx(2 ** -3.5) and loss.gammas[-1].item() == pytest.approx(2.0)
    assert loss.betas.sum().item() == pytest.approx(1.0, abs=1e-5)
    assert torch.allclose(torch.linalg.norm(loss.normalize(X), dim=1), torch.ones(11))


def test_qp_solver_matches_scipy() -> None:
    rng = np.random.default_rng(0)
    for trial in range(5):
        k = 8
        a = rng.normal(size=(k, k))
        q = a @ a.T + 0.1 * np.eye(k)
        d = rng.normal(size=k) + (0.5 if trial % 2 else 0.0)
        d[0] = abs(d[0]) + 0.1
        ours = solve_simplex_like_qp(torch.tensor(q), torch.tensor(d)).numpy()
        assert (ours >= -1e-9).all() and ours @ d == pytest.approx(1.0, abs=1e-8)
        res = optimize.minimize(lambda b: 0.5 * b @ q @ b, x0=np.where(d > 0, 1.0, 0.0) / max(d[d > 0].sum(), 1e-9), jac=lambda b: q @ b,
                                bounds=[(0, None)] * k, constraints=[{"type": "eq", "fun": lambda b: b @ d - 1, "jac": lambda b: d}],
                                method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
        assert 0.5 * ours @ q @ ours <= res.fun + 1e-7
    with pytest.raises(RuntimeError):
        solve_simplex_like_qp(torch.eye(3), torch.tensor([-1.0, -2.0, 0.0]))


@pytest.mark.parametrize("linear", [True, False])
def test_optimize_betas_non_degenerate(linear: bool) -> None:
    torch.manual_seed(42)
    loss = MkMmdLoss(DEVICE, perform_linear_approximation=linear)
    x = torch.randn(100, 5)
    y = (torch.randn(100, 5) + torch.tensor([1.0, 0, 0, 0, 0]) + torch.randn(100, 5) + torch.tensor([0, 1.0, 0, 0, 0])) / 2.0
    before = loss(x, y)
    h = loss.compute_all_h_u_linear(x, y) if linear else loss.compute_all_h_u_all_samples(x, y)
    hat_d = loss.compute_hat_d_per_kernel(h)
    assert before.item() == pytest.approx(float((loss.betas * hat_d).sum()), abs=1e-6)
    betas = loss.optimize_betas(x, y, 1e-4)
    assert betas.shape == (19, 1) and betas.sum().item() == pytest.approx(1.0, abs=1e-5) and bool((betas >= 0).all())
    if not linear:
        loss.betas = betas
        assert loss(x, y).item() > before.item()
    vertex = MkMmdLoss(DEVICE, minimize_type_two_error=False, perform_linear_approximation=linear).optimize_betas(x, y, 1e-4)
    assert int((vertex > 0).sum()) == 1 and vertex.sum().item() == pytest.approx(1.0)
    # degenerate: identical distributions with negative estimates fall back to a single kernel
    same = MkMmdLoss(DEVICE, perform_linear_approximation=linear)
    fallback = same.beta_with_extreme_kernel_base_values(-torch.ones(19, 1), torch.eye(19))
    assert int(fallback.sum()) == 1


def test_deep_mmd_trains_kernel_and_separates_distributions() -> None:
    torch.manual_seed(0)
    loss = DeepMmdLoss(DEVICE, input_size=6, hidden_size=8, output_size=4, lr=0.01, optimization_steps=3)
    x, y_same, y_far = torch.randn(40, 6), torch.randn(40, 6), torch.randn(40, 6) + 2.0
    weights_before = [p.detach().clone() for p in loss.featurizer.parameters()]
    value = loss(x, y_far)
    assert all(torch.equal(a, b) for a, b in zip(weights_before, loss.featurizer.parameters()))  # eval mode: no training
    assert value.item() > loss(x, y_same).item()
    loss.training = True
    x_req = x.clone().requires_grad_(True)
    out = loss(x_req, y_far)
    assert any(not torch.equal(a, b) for a, b in zip(weights_before, loss.featurizer.parameters()))
    out.backward()
    assert x_req.grad is not None and torch.isfinite(x_req.grad).all()
    biased = DeepMmdLoss(DEVICE, input_size=6, is_unbiased=False)
    assert torch.isfinite(biased(x, y_far))
