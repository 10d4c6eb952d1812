"""Pin the oracle's closed-form kernel functionals against a torch-autograd
restatement of the reference's nested-autodiff code (models/GP.py:28-180, 630-687)."""
import numpy as np
import pytest
import torch

from oracle.equation import EquationOracle
from oracle.gp import GPOracle, _Pairs
from tests.ref_autograd import RefKernels


@pytest.mark.parametrize("d", [6, 11])
def test_sixteen_functionals_match_autograd(d):
    rng = np.random.default_rng(d)
    idx = rng.choice(d, 5, replace=False)
    X = rng.uniform(-0.5, 0.5, (3, d + 1))
    Y = rng.uniform(-0.5, 0.5, (4, d + 1))
    ref = RefKernels(d, idx)
    a = 1.0 / (0.25 * np.sqrt(d)) ** 2
    P = _Pairs(X, Y, a, idx, d)
    for name, (rowop, colop) in RefKernels.TABLE.items():
        got = P.block(rowop, colop)
        fn = getattr(ref, name)
        for i in range(len(X)):
            for j in range(len(Y)):
                want = fn(torch.tensor(X[i], dtype=torch.float64), torch.tensor(Y[j], dtype=torch.float64)).item()
                assert got[i, j] == pytest.approx(want, rel=1e-11, abs=1e-13), (name, i, j)


def _toy_gp(d, n_dom, n_bdy, seed):
    rng = np.random.default_rng(seed)
    eq = EquationOracle(d + 1)
    gp = GPOracle(eq, idx_set=rng.choice(d, 5, replace=False))
    gp.x_t_domain = rng.uniform(-0.5, 0.5, (n_dom, d + 1))
    gp.x_t_boundary = rng.uniform(-0.5, 0.5, (n_bdy, d + 1))
    gp.N_domain, gp.N_boundary = n_dom, n_bdy
    gp.right_vector = rng.standard_normal((4 * n_dom + n_bdy, 1))
    return eq, gp, rng


def test_predict_and_gradient_match_autograd():
    d = 7
    eq, gp, rng = _toy_gp(d, 5, 3, 0)
    ref = RefKernels(d, gp.idx_set)
    xd = [torch.tensor(y, dtype=torch.float64) for y in gp.x_t_domain]
    xb = [torch.tensor(y, dtype=torch.float64) for y in gp.x_t_boundary]
    al = torch.tensor(gp.right_vector[:, 0], dtype=torch.float64)
    X = rng.uniform(-0.6, 0.6, (4, d + 1))
    u = gp.predict_raw(X)
    g = gp.gradient_raw(X)
    for i in range(len(X)):
        x = torch.tensor(X[i], dtype=torch.float64, requires_grad=True)
        val = ref.solution_function(x, xd, xb, al)
        (gx,) = torch.autograd.grad(val, x)
        assert u[i] == pytest.approx(val.item(), rel=1e-11, abs=1e-13)
        np.testing.assert_allclose(g[i], gx.numpy(), rtol=1e-10, atol=1e-12)


def test_pde_rows_match_autograd_rows():
    # compute_PDE_loss rows (models/GP.py:326-411) = row functionals @ alpha
    d = 6
    eq, gp, rng = _toy_gp(d, 4, 2, 1)
    ref = RefKernels(d, gp.idx_set)
    al = gp.right_vector[:, 0]
    X = rng.uniform(-0.5, 0.5, (2, d + 1))
    eps, u, dv, lp, dt = gp.pde_terms_raw(X)
    names = {
        "div": ["div_x_kappa", "div_x_kappa", "div_x_laplacian_y_t_kappa", "div_x_dt_y_t_kappa", "div_x_div_y_kappa"],
        "lap": ["laplacian_x_t_kappa", "laplacian_x_t_kappa", "laplacian_x_t_laplacian_y_t_kappa",
                "laplacian_x_t_dt_y_t_kappa", "laplacian_x_t_div_y_kappa"],
        "dt": ["dt_x_t_kappa", "dt_x_t_kappa", "dt_x_t_laplacian_y_t_kappa", "dt_x_t_dt_y_t_kappa", "dt_x_t_div_y_kappa"],
    }
    sets = [gp.x_t_domain, gp.x_t_boundary, gp.x_t_domain, gp.x_t_domain, gp.x_t_domain]
    got = {"div": dv, "lap": lp, "dt": dt}
    for op, fns in names.items():
        for i in range(len(X)):
            x = torch.tensor(X[i], dtype=torch.float64)
            row = []
            for fn, S in zip(fns, sets):
                row += [getattr(ref, fn)(x, torch.tensor(y, dtype=torch.float64)).item() for y in S]
            want = float(np.dot(np.array(row), al))
            assert got[op][i] == pytest.approx(want, rel=1e-10, abs=1e-12), (op, i)
    s = 0.25
    np.testing.assert_allclose(eps, dt + (s * s * u - 1 / d - s * s / 2) * dv + (s * s / 2) * lp, rtol=1e-13)


def test_gram_symmetric_and_layout():
    d = 5
    eq, gp, rng = _toy_gp(d, 6, 3, 2)
    K = gp.gram(gp.x_t_domain, gp.x_t_boundary, f16_entries=False)
    assert K.shape == (27, 27)
    np.testing.assert_allclose(K, K.T, rtol=0, atol=1e-13)
    ref = RefKernels(d, gp.idx_set)
    # spot-check one entry per off-diagonal block against the reference functionals (models/GP.py:196-248)
    xd, xb = gp.x_t_domain, gp.x_t_boundary
    T = lambda v: torch.tensor(v, dtype=torch.float64)
    N, Nb = 6, 3
    assert K[1, N + 2] == pytest.approx(ref.kappa(T(xd[1]), T(xb[2])).item(), rel=1e-12)                       # K12
    assert K[N + Nb + 1, 2] == pytest.approx(ref.laplacian_x_t_kappa(T(xd[1]), T(xd[2])).item(), rel=1e-11)     # K31
    assert K[2 * N + Nb + 3, N + Nb + 4] == pytest.approx(
        ref.dt_x_t_laplacian_y_t_kappa(T(xd[3]), T(xd[4])).item(), rel=1e-11)                                  # K43
    assert K[3 * N + Nb + 5, 2 * N + Nb] == pytest.approx(ref.div_x_dt_y_t_kappa(T(xd[5]), T(xd[0])).item(), rel=1e-11)  # K54
    assert K[N + 1, 3 * N + Nb + 2] == pytest.approx(ref.div_y_kappa(T(xb[1]), T(xd[2])).item(), rel=1e-11)     # K25
