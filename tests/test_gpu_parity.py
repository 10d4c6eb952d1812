"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped Python classes) against the
CPU oracle on identical inputs.  Tolerances: bit-exact for the sampler and integer bookkeeping;
1e-9 relative for FP64 kernels; the north-star criterion (rel-L2 / L1 error vs the exact solution within
1e-6 relative of the oracle's) for the solvers."""
import copy
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import rng as orng
from oracle.equation import EquationOracle
from oracle.gp import GPOracle
from oracle.solvers import (MLPFullHistoryOracle, MLPOracle, ScaSMLFullHistoryOracle, ScaSMLOracle)


@pytest.fixture(autouse=True)
def _fp64_route_for_this_module():
    """This module checks the FP64 parity-anchor route against the oracle at 1e-9..1e-10; the solvers' default for the
    sampled points is the tcgen05 route, which has its own (north-star tolerance) tests in test_gpu_tc.py."""
    from scasml_gp_b200 import _lib
    from scasml_gp_b200.solvers._picard import PicardSolverBase
    old = PicardSolverBase.route
    PicardSolverBase.route = _lib.ROUTE_F64
    yield
    PicardSolverBase.route = old


def _product():
    import torch
    from scasml_gp_b200 import _lib
    from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
    from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
    from scasml_gp_b200.solvers.MLP import MLP
    from scasml_gp_b200.solvers.MLP_full_history import MLP_full_history
    from scasml_gp_b200.solvers.ScaSML import ScaSML
    from scasml_gp_b200.solvers.ScaSML_full_history import ScaSML_full_history
    return dict(torch=torch, lib=_lib, Eq=Grad_Dependent_Nonlinear, GP=GP_Grad_Dependent_Nonlinear, MLP=MLP,
                MLPfh=MLP_full_history, ScaSML=ScaSML, ScaSMLfh=ScaSML_full_history)


class Fitted:
    """Oracle GP and product GP fitted on the same collocation points / index set / Newton start."""

    def __init__(self, d, nd, nb, seed=1234):
        P = _product()
        self.P = P
        self.d = d
        self.eq_o = EquationOracle(d + 1)
        self.dom, self.bdy = self.eq_o.generate_data(nd, nb, seed=seed)
        self.idx = np.random.default_rng(7).choice(d, 5, replace=False)
        self.sol0 = np.random.default_rng(0).standard_normal(3 * nd) * 1e-3
        self.gp_o = GPOracle(self.eq_o, idx_set=self.idx)
        self.gp_o.GPsolver(self.dom, self.bdy, sol0=self.sol0)
        self.eq = P["Eq"](d + 1)
        self.gp = P["GP"](self.eq, idx_set=self.idx)
        self.gp.GPsolver(self.dom, self.bdy, sol0=self.sol0)

    def test_points(self, n_dom, n_bdy, seed=42):
        return np.concatenate(self.eq_o.generate_test_data(n_dom, n_bdy, seed=seed), axis=0)


@pytest.fixture(scope="module")
def small():
    return Fitted(d=6, nd=40, nb=12)


@pytest.fixture(scope="module")
def mid():
    return Fitted(d=20, nd=200, nb=40)


def test_sampler_bit_exact():
    P = _product()
    torch, lib = P["torch"], P["lib"]
    lib.ensure_normal_table(debug=True)                 # the draw hook lives in the debug build of the library (same sampler code)
    assert np.array_equal(lib.normal_half_table().view(np.uint16), orng.normal_half_table().view(np.uint16))
    for (stream, domain, seed, start, count) in [(0, 0, 0, 0, 4099), (17, 1, 0, 123457, 3001), (3, 1, 9, 2**33 + 5, 1000)]:
        out = torch.empty(count, dtype=torch.float64, device="cuda")
        for uniform in (0, 1):
            lib.check(lib.load_debug().scasml_debug_draw(stream, domain, seed, start, count, uniform, lib.ptr(out), lib.stream_ptr()))
            key = orng.make_key(stream, domain, seed)
            want = orng.uniforms(key, start, count) if uniform else orng.normals(key, start, count)
            assert np.array_equal(out.cpu().numpy(), want)


def test_equation_f_g(small):
    X = small.test_points(33, 7)
    np.testing.assert_array_equal(small.eq.g(X), small.eq_o.g(X).astype(np.float16))
    np.testing.assert_array_equal(small.eq.exact_solution(X), small.eq_o.exact_solution(X).astype(np.float16))
    rng = np.random.default_rng(0)
    u, z = rng.standard_normal((40, 1)), rng.standard_normal((40, small.d))
    got = small.eq.f(X, u, z).astype(np.float64)
    want = small.eq_o.f(X, u, z).astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-7)          # float16 outputs, 1-ulp flips allowed
    assert small.eq.mu() == small.eq_o.mu() and small.eq.sigma() == small.eq_o.sigma()


def test_gram_matches_oracle(small):
    P = small.P
    torch, lib = P["torch"], P["lib"]
    gp = small.gp
    K = torch.empty((gp.phi_dim, gp.phi_dim), dtype=torch.float64, device="cuda")
    lib.check(lib.load().scasml_gp_gram(gp._handle, lib.ptr(K), 0, 0, lib.stream_ptr()))
    want = small.gp_o.gram(small.dom, small.bdy, f16_entries=False)
    np.testing.assert_allclose(K.cpu().numpy(), want, rtol=1e-11, atol=1e-13)
    Kp = gp.kernel_phi_phi(small.dom, small.bdy).astype(np.float64)          # float16, nugget on the diagonal
    want16 = (small.gp_o.gram(small.dom, small.bdy) + 1e-2 * np.eye(gp.phi_dim)).astype(np.float16).astype(np.float64)
    assert np.mean(Kp != want16) < 1e-4
    gp.GPsolver(small.dom, small.bdy, sol0=small.sol0)                       # kernel_phi_phi rebinds; refit


def test_dense_pieces(small):
    P = small.P
    torch, lib = P["torch"], P["lib"]
    rng = np.random.default_rng(3)
    for n in (37, 200, 321):
        A = rng.standard_normal((n, n))
        S = A @ A.T + n * np.eye(n)
        Sd = torch.from_numpy(S.copy()).cuda()
        Pd = torch.empty((n, n), dtype=torch.float64, device="cuda")
        ws = torch.empty((n * n + 64 * 64 * ((n + 63) // 64) + 64 * n) * 8 + 4096, dtype=torch.uint8, device="cuda")
        lib.check(lib.load_debug().scasml_debug_spd_inverse(lib.ptr(Sd), n, lib.ptr(Pd), lib.ptr(ws), ws.numel(), lib.stream_ptr()))
        np.testing.assert_allclose(Pd.cpu().numpy(), np.linalg.inv(S), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.tril(Sd.cpu().numpy()), np.linalg.cholesky(S), rtol=1e-10, atol=1e-12)
        H = rng.standard_normal((n, n)) + 0.1 * np.eye(n)                    # indefinite, needs pivoting
        b = rng.standard_normal(n)
        Hd, bd = torch.from_numpy(H.copy()).cuda(), torch.from_numpy(b.copy()).cuda()
        lib.check(lib.load_debug().scasml_debug_lu_solve(lib.ptr(Hd), n, lib.ptr(bd), lib.stream_ptr()))
        want = np.linalg.solve(H, b)
        np.testing.assert_allclose(bd.cpu().numpy(), want, rtol=1e-7, atol=1e-9 * np.abs(want).max())


def test_fit_matches_oracle(small, mid):
    for F in (small, mid):
        a_o = F.gp_o.right_vector[:, 0]
        a_p = F.gp.right_vector[:, 0]
        assert np.linalg.norm(a_p - a_o) / np.linalg.norm(a_o) < 1e-7
        assert F.gp.newton_steps == F.gp_o.newton_steps
        np.testing.assert_allclose(F.gp.loss_history, F.gp_o.loss_history, rtol=1e-8)
        got = F.gp.predict(F.dom).astype(np.float64)
        want = F.gp_o.predict(F.dom).astype(np.float64)
        assert np.mean(got != want) < 0.01


def test_eval_modes_match_oracle(small, mid):
    for F in (small, mid):
        P = F.P
        lib = P["lib"]
        gp = F.gp
        gp.set_right_vector(F.gp_o.right_vector)            # identical weights -> isolates the evaluation kernels
        X = F.test_points(150, 31)
        X[:5] += 0.37                                        # off-grid (non-float16) coordinates
        np.testing.assert_allclose(gp.predict_raw(X), F.gp_o.predict_raw(X), rtol=1e-10, atol=1e-12)
        eps_o, u_o, dv_o, lp_o, dt_o = F.gp_o.pde_terms_raw(X)
        eps, dv, lp, dt = gp.pde_terms_raw(X)
        for a, b in ((eps, eps_o), (dv, dv_o), (lp, lp_o), (dt, dt_o)):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-10 * np.abs(b).max())
        u, G = gp._eval(X, lib.EVAL_UG, nout=2)
        np.testing.assert_allclose(u.cpu().numpy(), u_o, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(G.cpu().numpy(), F.gp_o.gradient_raw(X)[:, :-1].sum(1), rtol=1e-9, atol=1e-10)
        (gt,) = gp._eval(X, lib.EVAL_TERMINAL)
        np.testing.assert_allclose(gt.cpu().numpy(), F.eq_o.g(X, cast=False)[:, 0] - u_o, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(gp.gradient_raw(X), F.gp_o.gradient_raw(X), rtol=1e-9, atol=1e-10)
        assert gp.predict(X).dtype == np.float16 and gp.predict(X).shape == (len(X), 1)
        assert gp.compute_gradient(X, None).shape == (len(X), F.d + 1)
        assert gp.compute_PDE_loss(X).shape == (len(X), 1)
        assert gp.predict(X[:0]).shape == (0, 1)             # empty batch


def _rel_l2(sol, exact):
    sol, exact = np.asarray(sol, dtype=np.float64).ravel(), np.asarray(exact, dtype=np.float64).ravel()
    m = ~(np.isnan(sol) | np.isnan(exact))                    # tests/SimpleUniform.py:110-136
    return np.linalg.norm(sol[m] - exact[m]) / np.linalg.norm(exact[m]), np.mean(np.abs(sol[m] - exact[m]))


def _check_solver(prod, orac, args_p, args_o, X, exact):
    got = prod.u_solve(*args_p)
    want = orac.u_solve(*args_o)
    assert got.dtype == np.float16 and got.shape == (len(X), 1)
    np.testing.assert_allclose(prod.last_raw, orac.last_raw, rtol=1e-8, atol=1e-11)
    assert prod.evaluation_counter == orac.evaluation_counter
    assert prod.key == orac.key_counter
    l2_p, l1_p = _rel_l2(prod.last_raw_u, exact)
    l2_o, l1_o = _rel_l2(orac.last_raw_u, exact)
    assert abs(l2_p - l2_o) <= 1e-6 * l2_o and abs(l1_p - l1_o) <= 1e-6 * l1_o        # north-star criterion
    assert np.mean(got.astype(np.float64) != want.astype(np.float64)) < 0.02             # float16 flips only


@pytest.mark.parametrize("n", [1, 2, 3])
def test_scasml_quadrature_parity(small, n):
    F, P = small, small.P
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(21, 6)
    exact = F.eq_o.exact_solution(X)
    prod = P["ScaSML"](F.eq, F.gp)
    orac = ScaSMLOracle(F.eq_o, F.gp_o, cast=False)
    if n == 1:                                  # the reference's own rho=1 table has a NaN weight
        prod.quadrature = "gauss_legendre"
        orac.true_gl = True
    _check_solver(prod, orac, (n, n, X), (n, n, X), X, exact)
    # second call: the split counter persists (solvers/ScaSML.py:27,228) -> different step normals
    first = prod.last_raw.copy()
    _check_solver(prod, orac, (n, n, X), (n, n, X), X, exact)
    assert not np.array_equal(first, prod.last_raw)


def test_degenerate_steps_fall_back_to_regenerated_normals(small):
    """The reduction recovers the Brownian increments from the sampled points; when a point equals its parent (test point
    at the terminal time T, so every step has zero length) it must regenerate the Philox normals instead: the terminal
    z = mean(g N) / (T - t + 1e-6) of solvers/ScaSML.py:211-215 and the full-history y N / sqrt(step + 1e-6) still need N."""
    F, P = small, small.P
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(14, 4)
    X[::2, -1] = F.eq_o.T                                  # every other test point sits at the terminal time
    exact = F.eq_o.exact_solution(X)
    prod, orac = P["ScaSML"](F.eq, F.gp), ScaSMLOracle(F.eq_o, F.gp_o, cast=False)
    prod.quadrature = "gauss_legendre"
    orac.true_gl = True
    prod.uz_solve(2, 2, X)
    orac.uz_solve(2, 2, X)
    np.testing.assert_allclose(prod.last_raw, orac.last_raw, rtol=1e-8, atol=1e-11)
    assert np.all(np.isfinite(prod.last_raw)) and np.abs(prod.last_raw[::2, 1:]).max() > 0
    prod, orac = P["ScaSMLfh"](F.eq, F.gp), ScaSMLFullHistoryOracle(F.eq_o, F.gp_o, cast=False)
    prod.uz_solve(2, None, X, 3)
    orac.uz_solve(2, None, X, M=3)
    np.testing.assert_allclose(prod.last_raw, orac.last_raw, rtol=1e-8, atol=1e-11)
    prod, orac = P["MLP"](F.eq), MLPOracle(F.eq_o, cast=False)
    prod.quadrature = "gauss_legendre"
    orac.true_gl = True
    prod.uz_solve(2, 2, X)
    orac.uz_solve(2, 2, X)
    np.testing.assert_allclose(prod.last_raw, orac.last_raw, rtol=1e-8, atol=1e-11)


def test_scasml_reference_tables_nan_semantics(small):
    F, P = small, small.P
    X = F.test_points(9, 2)
    prod = P["ScaSML"](F.eq, F.gp)
    orac = ScaSMLOracle(F.eq_o, F.gp_o, cast=False)
    got = prod.u_solve(1, 1, X)                 # NaN weight in the N=2 table -> NaN everywhere
    want = orac.u_solve(1, 1, X)
    assert np.all(np.isnan(got.astype(np.float64))) and np.all(np.isnan(want.astype(np.float64)))


@pytest.mark.parametrize("n", [1, 2, 3])
def test_mlp_quadrature_parity(small, n):
    F, P = small, small.P
    X = F.test_points(25, 5)
    exact = F.eq_o.exact_solution(X)
    prod, orac = P["MLP"](F.eq), MLPOracle(F.eq_o, cast=False)
    if n == 1:
        prod.quadrature = "gauss_legendre"
        orac.true_gl = True
    _check_solver(prod, orac, (n, n, X), (n, n, X), X, exact)


@pytest.mark.parametrize("n,M", [(1, 3), (2, 3), (3, 2), (1, 7)])
def test_full_history_parity(small, n, M):
    F, P = small, small.P
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(17, 4)
    exact = F.eq_o.exact_solution(X)
    _check_solver(P["ScaSMLfh"](F.eq, F.gp), ScaSMLFullHistoryOracle(F.eq_o, F.gp_o, cast=False),
                  (n, None, X, M), (n, None, X, M), X, exact)
    _check_solver(P["MLPfh"](F.eq), MLPFullHistoryOracle(F.eq_o, cast=False), (n, None, X, M), (n, None, X, M), X, exact)


def test_level_zero_and_empty_batch(small):
    F, P = small, small.P
    X = F.test_points(5, 1)
    s = P["ScaSML"](F.eq, F.gp)
    uz = s.uz_solve(0, 2, X)
    assert uz.shape == (6, F.d + 1) and not uz.any()
    assert s.evaluation_counter == 2            # g() call + MC_g = 1 (solvers/ScaSML.py:59,205)
    assert s.uz_solve(2, 2, X[:0]).shape == (0, F.d + 1)


def test_batching_is_invisible(mid):
    F, P = mid, mid.P
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(40, 8)
    a = P["ScaSML"](F.eq, F.gp)
    a.u_solve(2, 2, X)
    b = P["ScaSML"](F.eq, F.gp)
    ws1, _ = b.plan(2, 2, 1)
    b.workspace_budget_bytes = 7 * ws1          # forces 7-row chunks with a ragged tail
    b.u_solve(2, 2, X)
    assert np.array_equal(a.last_raw, b.last_raw)
    assert a.evaluation_counter == b.evaluation_counter and a.key == b.key


def test_sample_sharding_partials_sum_to_full(mid):
    """ranks own the units u = rank + world*s of every top-level sample array; the un-clipped weighted
    partial sums add up to the unsharded result.  Emulated on one GPU by running every rank's shard in turn."""
    F, P = mid, mid.P
    lib = P["lib"]
    torch = P["torch"]
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(10, 3)
    xd = lib.to_device(X)
    for cls, args in ((P["ScaSML"], dict(n=3, rho=3, M=None)), (P["ScaSMLfh"], dict(n=2, rho=None, M=3))):
        want = cls(F.eq, F.gp)._uz_device(args["n"], args["rho"], xd, args["M"]).cpu().numpy()
        for world in (2, 3, 8):
            acc = torch.zeros((len(X), F.d + 1), dtype=torch.float64, device="cuda")
            for rank in range(world):
                s = cls(F.eq, F.gp)
                p = s._params(args["n"], args["rho"], args["M"], rank, world)
                need = C.c_size_t(0)
                lib.check(lib.load().scasml_picard_plan(C.byref(p), len(X), C.byref(need), None))
                ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
                out = torch.empty_like(acc)
                st = lib.PicardStats()
                lib.check(lib.load().scasml_uz_solve(F.gp._handle, C.byref(p), 0, lib.ptr(xd), len(X), lib.ptr(out),
                                                     lib.ptr(ws), ws.numel(), C.byref(st), lib.stream_ptr()))
                acc += out
            lib.check(lib.load().scasml_clip(lib.ptr(acc), acc.numel(), 0.1, lib.stream_ptr()))
            np.testing.assert_allclose(acc.cpu().numpy(), want, rtol=1e-9, atol=1e-13)


def test_deepcopy_and_refit(small):
    F, P = small, small.P
    X = F.test_points(12, 3)
    F.gp.set_right_vector(F.gp_o.right_vector)
    s = P["ScaSML"](F.eq, F.gp)
    s2 = copy.deepcopy(s)                        # tests/ComputingBudget.py:138
    assert s2.GP is not s.GP and s2.GP._handle.value != s.GP._handle.value
    a, b = s.u_solve(2, 2, X), s2.u_solve(2, 2, X)
    assert np.array_equal(a, b)
    s2.GP.GPsolver(F.dom[:20], F.bdy[:6], GN_steps=5)      # in-place refit of the copy (tests/ComputingBudget.py:171)
    assert s2.GP.N_domain == 20 and s.GP.N_domain == 40
    assert np.array_equal(s.GP.predict(X), F.gp.predict(X))
    s.PINN = s.GP                                # arbitrary attribute assignment (tests/RepeatedExperiment.py:180)


def test_unfitted_gp_and_missing_inputs_fail_loudly(small):
    P = small.P
    gp = P["GP"](small.eq)
    with pytest.raises(AttributeError):
        gp.predict(small.test_points(3, 1))
    with pytest.raises(ValueError):
        P["ScaSML"](small.eq, small.gp).u_solve(3, 2, small.test_points(3, 1))     # n > rho


def test_config2_size_properties():
    """BASELINE config 2 (d=20, n=rho=3, 1000+200 collocation points): oracle parity on a few test points, plus
    size-independent properties on the full 1 200-point batch (batching invariance, finite outputs, clip bound)."""
    F = Fitted(d=20, nd=1000, nb=200)
    P = F.P
    assert np.linalg.norm(F.gp.right_vector - F.gp_o.right_vector) / np.linalg.norm(F.gp_o.right_vector) < 1e-6
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(1000, 200)
    exact = F.eq_o.exact_solution(X)
    prod = P["ScaSML"](F.eq, F.gp)
    u = prod.u_solve(3, 3, X).astype(np.float64)
    raw_full = prod.last_raw.copy()
    assert np.all(np.abs(raw_full[~np.isnan(raw_full)]) <= 0.1 + 1e-12)
    assert prod.evaluation_counter == 549 and prod.last_stats["sample_points"] == 1090
    l2_gp, _ = _rel_l2(F.gp.predict(X), exact)
    l2_sc, _ = _rel_l2(u, exact)
    assert l2_sc < l2_gp                                         # the correction helps
    # oracle parity on a slice, addressed by global row ids
    sl = slice(100, 106)
    orac = ScaSMLOracle(F.eq_o, F.gp_o, cast=False)
    orac.uz_solve(3, 3, X[sl], gid0=100)
    np.testing.assert_allclose(raw_full[sl], orac.last_raw, rtol=1e-7, atol=1e-10)
    # batching invariance at full size
    b = P["ScaSML"](F.eq, F.gp)
    ws1, _ = b.plan(3, 3, 1)
    b.workspace_budget_bytes = 500 * ws1
    b.u_solve(3, 3, X)
    assert np.array_equal(np.nan_to_num(b.last_raw, nan=7.0), np.nan_to_num(raw_full, nan=7.0))


@pytest.mark.parametrize("d", [33, 150, 1100])
def test_sampler_column_pass_variants_through_the_mlp_solvers(d):
    """The samplers are templated on the number of 32-column passes over a point (d + 1 <= 32, 64, 128 keep the point in registers;
    d + 1 <= 256, 1024, 2048 walk it in load rounds): the surrogate-free MLP solvers exercise them at any d against the oracle
    (quadrature: terminal + chained path kernels; full history: uniform time draws)."""
    P = _product()
    eq_o = EquationOracle(d + 1)
    eq = P["Eq"](d + 1)
    X = np.concatenate(eq_o.generate_test_data(9, 3, seed=5), axis=0)
    X[3, -1] = eq_o.T                                       # a degenerate row (T == t)
    exact = eq_o.exact_solution(X)
    prod, orac = P["MLP"](eq), MLPOracle(eq_o, cast=False)
    _check_solver(prod, orac, (2, 2, X), (2, 2, X), X, exact)
    prod, orac = P["MLPfh"](eq), MLPFullHistoryOracle(eq_o, cast=False)
    _check_solver(prod, orac, (2, None, X, 3), (2, None, X, 3), X, exact)


def _abi_comm_worker(rank, world, idfile, q):
    import os, time
    import torch
    torch.cuda.set_device(rank)
    P = _product()
    lib = P["lib"]
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(lib.AbiComm.unique_id())
        os.replace(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.01)
    comm = lib.AbiComm(open(idfile, "rb").read(), rank, world)
    F = Fitted(d=6, nd=40, nb=12)
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(9, 3)
    one = P["ScaSML"](F.eq, F.gp)
    one.u_solve(2, 2, X)
    many = P["ScaSML"](F.eq, F.gp)
    many.comm = comm
    many.u_solve(2, 2, X)
    diff = float(np.nanmax(np.abs(many.last_raw - one.last_raw)))
    same = bool(np.array_equal(np.isnan(many.last_raw), np.isnan(one.last_raw))) and many.evaluation_counter == one.evaluation_counter
    torch.cuda.synchronize()
    comm.close()
    q.put((rank, diff, same))


def test_allreduce_through_the_c_abi_reproduces_the_single_gpu_solve(tmp_path):
    """scasml_comm_* (NCCL loaded by the library): two processes, one GPU each, the sample units of a level-2 solve sharded over them and
    the partial (u, z) blocks summed by scasml_allreduce_partial -- no torch.distributed."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = str(tmp_path / "nccl_id")
    procs = [ctx.Process(target=_abi_comm_worker, args=(r, 2, idfile, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, diff, same in res:
        assert same and diff < 1e-12, (rank, diff, same)
