"""Known-answer tests for the oracle's recursions (integer tables, call counts,
evaluation counters confirmed by the reference's committed profiles / InferenceScaling plots)."""
import numpy as np
import pytest

from oracle import rng as orng
from oracle.equation import EquationOracle
from oracle.gp import GPOracle
from oracle.solvers import (MLPFullHistoryOracle, MLPOracle, ScaSMLFullHistoryOracle, ScaSMLOracle)
from oracle.tables import approx_parameters


class StubGP:
    def predict_raw(self, X):
        return np.zeros(len(X))

    def gradient_raw(self, X):
        return np.zeros((len(X), X.shape[1]))

    def pde_raw(self, X):
        return np.zeros(len(X))


def _pts(d, B, seed=0):
    eq = EquationOracle(d + 1)
    dom, bdy = eq.generate_test_data(B, 0, seed=seed)
    return eq, dom


def test_tables_app_c1():
    want_Q = {1: [2], 2: [3, 3], 3: [3, 3, 4], 4: [3, 4, 4, 5], 5: [3, 4, 4, 5, 6]}
    want_Mf = {1: [1], 2: [1, 2], 3: [2, 3, 5], 4: [2, 4, 8, 16], 5: [2, 5, 11, 25, 56]}
    want_Mg = {1: [1, 1], 2: [1, 2, 4], 3: [1, 3, 9, 27], 4: [1, 4, 16, 64, 256], 5: [1, 5, 25, 125, 625, 3125]}
    for rho in range(1, 6):
        Mf, Mg, Q, c, w = approx_parameters(rho)
        assert list(Q[rho - 1, :rho]) == want_Q[rho]
        assert list(Mf[rho - 1, :rho]) == want_Mf[rho]
        assert list(Mg[rho - 1, :rho + 1]) == want_Mg[rho]


def test_lgwt_defect_app_c2():
    Mf, Mg, Q, c, w = approx_parameters(4)
    assert c[0, 0] == 0.25 and w[0, 0] == 0.5
    assert np.isnan(w[0, 1]) and c[0, 1] == 0.0 and c[1, 1] == pytest.approx(0.394338, abs=1e-6)
    np.testing.assert_allclose(c[:3, 2], [0.16393370341761293, 0.16393370341761299, 0.44364916731037085], rtol=0, atol=2e-16)
    np.testing.assert_allclose(w[:3, 2], [0.060458, 0.060458, 0.138889], atol=1e-6)
    assert w[:3, 2].sum() == pytest.approx(0.2598, abs=1e-4)          # not 0.5: not Gauss-Legendre
    np.testing.assert_allclose(c[:4, 3], [0.0871313, 0.2943559497716291, 0.2943559497716292, 0.4652841], atol=1e-7)
    assert w[:4, 3].sum() == pytest.approx(0.2515, abs=1e-4)
    assert c[1, 4] < c[0, 4]                                          # N=5: second step negative -> sqrt NaN
    Mf, Mg, Q, c, w = approx_parameters(4, true_gl=True)
    for q in range(1, 6):
        assert w[:q, q - 1].sum() == pytest.approx(0.5, rel=1e-13)
        assert np.all(np.diff(c[:q, q - 1]) > 0)


@pytest.mark.parametrize("n,counter,sp", [(1, 10, None), (2, 83, 46), (3, 549, 1090), (4, 3659, 13426)])
def test_scasml_counter_and_sample_points(n, counter, sp):
    eq, X = _pts(4, 1)
    s = ScaSMLOracle(eq, StubGP(), true_gl=True)
    s.u_solve(n, n, X)
    assert s.evaluation_counter == counter
    if sp is not None:
        assert s.sample_points == sp


@pytest.mark.parametrize("n,counter", [(1, 5), (2, 46), (3, 372), (4, 2714)])
def test_mlp_counter(n, counter):
    eq, X = _pts(4, 1)
    s = MLPOracle(eq, true_gl=True)
    s.u_solve(n, n, X)
    assert s.evaluation_counter == counter


@pytest.mark.parametrize("n,c_sca,c_mlp,sp", [(1, 10, 7, 9), (2, 54, 33, 60), (3, 246, 127, 390), (4, 1034, 449, 2523)])
def test_full_history_counters(n, c_sca, c_mlp, sp):
    eq, X = _pts(4, 1)
    s = ScaSMLFullHistoryOracle(eq, StubGP())
    s.u_solve(n, None, X, M=3)
    assert s.evaluation_counter == c_sca and s.sample_points == sp
    m = MLPFullHistoryOracle(eq)
    m.u_solve(n, None, X, M=3)
    assert m.evaluation_counter == c_mlp


def test_counter_is_cumulative_like_inference_scaling():
    # tests/InferenceScaling.py:142 reads the running counter: 10, 64, 310 in the committed plots
    eq, X = _pts(4, 1)
    s = ScaSMLFullHistoryOracle(eq, StubGP())
    seen = []
    for n in (1, 2, 3):
        s.u_solve(n, None, X, M=3)
        seen.append(s.evaluation_counter)
    assert seen == [10, 64, 310]


def test_rng_aliasing_structure():
    # value depends only on (key, flat index); same key + same size -> same stream
    k = orng.make_key(0, orng.DOMAIN_FIXED)
    a = orng.normals(k, 0, 2 * 9 * 5).reshape(2, 9, 5)
    b = orng.normals(k, 0, 6 * 3 * 5).reshape(6, 3, 5)
    assert np.array_equal(a.reshape(-1), b.reshape(-1))
    assert np.array_equal(orng.normals(k, 40, 50), a.reshape(-1)[40:])
    tau = orng.uniforms(k, 0, 64)
    n = orng.normals(k, 0, 64)
    assert np.all(np.diff(tau[np.argsort(n, kind="stable")]) >= 0)     # tau monotone in the normal (tau ~ Phi(N))
    assert np.array_equal(n.astype(np.float16).astype(np.float64), n)  # float16-valued
    big = orng.normals(orng.make_key(7, orng.DOMAIN_STEP), 3, 400000)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3


def _small_fit(d=6, nd=40, nb=12, cast=True):
    eq = EquationOracle(d + 1)
    gp = GPOracle(eq, cast=cast)
    dom, bdy = eq.generate_data(nd, nb, seed=1234)
    gp.GPsolver(dom, bdy, GN_steps=20)
    return eq, gp


def test_fit_converges_and_satisfies_constraints():
    eq, gp = _small_fit()
    assert gp.loss_history[-1] < gp.loss_history[0]
    assert gp.newton_steps < 20                                       # clean arithmetic stops early
    # the fitted GP reproduces its own collocation data up to the nugget
    K = gp.gram(gp.x_t_domain, gp.x_t_boundary)
    z = gp._b(gp.sol)
    np.testing.assert_allclose((K + gp.nugget * np.eye(gp.phi_dim)) @ gp.right_vector[:, 0], z, atol=1e-8)


def test_sharded_partials_sum_to_unsharded():
    eq, gp = _small_fit()
    X = eq.generate_test_data(5, 2, seed=42)
    X = np.concatenate(X, axis=0)
    for cls, args in ((ScaSMLOracle, (2, 2)), (ScaSMLFullHistoryOracle, (2, None))):
        full = cls(eq, gp, cast=False)
        kw = {"M": 3} if cls is ScaSMLFullHistoryOracle else {}
        full.uz_solve(*args, X, **kw)
        world = 3
        acc = 0
        for r in range(world):
            part = cls(eq, gp, cast=False)
            acc = acc + part.uz_solve(*args, X, shard=(r, world), **kw)
        np.testing.assert_allclose(acc, full.partial_uz, rtol=1e-12, atol=1e-15)


def test_batching_invariance_via_global_row_ids():
    eq, gp = _small_fit()
    X = np.concatenate(eq.generate_test_data(6, 2, seed=43), axis=0)
    a = ScaSMLOracle(eq, gp, cast=False)
    full = a.uz_solve(2, 2, X)
    b = ScaSMLOracle(eq, gp, cast=False)
    lo = b.uz_solve(2, 2, X[:3], gid0=0)
    b.key_counter = 0
    hi = b.uz_solve(2, 2, X[3:], gid0=3)
    assert np.array_equal(np.concatenate([lo, hi]), full)


def test_scasml_improves_on_gp_small_case():
    # statistical sanity (the reference's headline claim): correction reduces the GP error
    eq, gp = _small_fit(d=10, nd=100, nb=20)
    X = np.concatenate(eq.generate_test_data(60, 12, seed=42), axis=0)
    exact = eq.exact_solution(X)[:, 0]
    rel = lambda s: np.linalg.norm(s - exact) / np.linalg.norm(exact)
    e_gp = rel(gp.predict(X)[:, 0].astype(np.float64))
    e_sc = rel(ScaSMLOracle(eq, gp).u_solve(2, 2, X)[:, 0].astype(np.float64))
    e_sc_nocast = rel(ScaSMLOracle(eq, gp, cast=False).u_solve(2, 2, X)[:, 0].astype(np.float64))
    assert e_sc < e_gp
    assert abs(e_sc - e_sc_nocast) < 0.1 * e_sc                       # cast vs nocast: statistically irrelevant
