"""GPU tests of the tcgen05 route: plumbing self-test (descriptors / swizzle / TMEM), evaluation parity against the
oracle, and the solver-level north-star criterion under the "nocast" policy."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.solvers import ScaSMLFullHistoryOracle, ScaSMLOracle
from tests.test_gpu_parity import Fitted, _product, _rel_l2


def _tc_gemm(lib, torch, A, B, lbo=1, sbo=64, layout=2, kstep=32):
    K, N = A.shape[1], B.shape[0]
    Ad = torch.from_numpy(A).cuda()
    Bd = torch.from_numpy(B).cuda()
    Dd = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    lib.check(lib.load_debug().scasml_debug_tc_gemm(lib.ptr(Ad), lib.ptr(Bd), lib.ptr(Dd), K, N, lbo, sbo, layout, kstep, lib.stream_ptr()))
    torch.cuda.synchronize()
    return Dd.cpu().numpy()


@pytest.mark.parametrize("K,N", [(64, 64), (128, 64), (128, 32), (256, 16)])
def test_tcgen05_plumbing_selftest(K, N):
    P = _product()
    torch, lib = P["torch"], P["lib"]
    rng = np.random.default_rng(K + N)
    A = rng.standard_normal((128, K)).astype(np.float16)
    B = rng.standard_normal((N, K)).astype(np.float16)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = _tc_gemm(lib, torch, A, B)
    err = np.max(np.abs(got - want))
    if not err < 1e-3:
        # diagnostics: which descriptor variant (if any) reproduces the product
        report = []
        for lbo in (0, 1, 64):
            for sbo in (64, 8, 128):
                for kstep in (32, 2):
                    g = _tc_gemm(lib, torch, A, B, lbo, sbo, 2, kstep)
                    report.append((lbo, sbo, kstep, float(np.nanmax(np.abs(g - want)))))
        pytest.fail(f"tcgen05 self-test max err {err}; variants (lbo, sbo, kstep, err): {report}")


@pytest.fixture(scope="module")
def mid():
    return Fitted(d=20, nd=200, nb=40)


@pytest.fixture(scope="module")
def wide():
    return Fitted(d=100, nd=300, nb=60)


def _eval_both(F, X):
    lib = F.P["lib"]
    gp = F.gp
    gp.set_right_vector(F.gp_o.right_vector)
    out = {}
    for route in (lib.ROUTE_F64, lib.ROUTE_TC):
        gp.route = route
        u = gp.predict_raw(X)
        eps, dv, lp, dt = gp.pde_terms_raw(X)
        uu, G = [t.cpu().numpy() for t in gp._eval(X, lib.EVAL_UG, nout=2)]
        (gt,) = gp._eval(X, lib.EVAL_TERMINAL)
        out[route] = dict(u=u, eps=eps, dv=dv, lp=lp, dt=dt, uu=uu, G=G, gt=gt.cpu().numpy())
    gp.route = lib.ROUTE_F64
    return out


@pytest.fixture(scope="module")
def huge():
    return Fitted(d=300, nd=150, nb=40)      # d + 2 > 128: K-streamed kernel (5 K blocks, partial last block)


@pytest.mark.parametrize("which", ["mid", "wide", "huge"])
def test_tc_eval_matches_fp64_route_and_oracle(which, request):
    F = request.getfixturevalue(which)
    lib = F.P["lib"]
    assert lib.load().scasml_gp_tc_supported(F.gp._handle) == 1
    X = F.test_points(301, 40)
    X[:7] += 0.123456789                                   # not float16-representable
    o = _eval_both(F, X)
    f64, tcr = o[lib.ROUTE_F64], o[lib.ROUTE_TC]
    np.testing.assert_allclose(f64["u"], F.gp_o.predict_raw(X), rtol=1e-10, atol=1e-12)
    scale = {k: max(1.0, float(np.abs(f64[k]).max())) for k in f64}
    for k in ("u", "uu", "gt"):
        assert np.max(np.abs(tcr[k] - f64[k])) < 2e-6 * scale[k], k
    for k in ("G", "dv", "lp", "dt", "eps"):
        assert np.max(np.abs(tcr[k] - f64[k])) < 2e-5 * scale[k], k
    # typical (rms) error is what enters the Monte-Carlo means
    assert np.sqrt(np.mean((tcr["u"] - f64["u"]) ** 2)) < 5e-7


def test_tc_solver_meets_north_star_criterion(mid):
    F, P = mid, mid.P
    lib = P["lib"]
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(48, 12)
    exact = F.eq_o.exact_solution(X)
    for cls, ocls, args, kw in ((P["ScaSML"], ScaSMLOracle, (3, 3), {}), (P["ScaSMLfh"], ScaSMLFullHistoryOracle, (3, None), {"M": 3})):
        prod = cls(F.eq, F.gp)
        prod.route = lib.ROUTE_TC
        F.gp.route = lib.ROUTE_TC
        prod.u_solve(*args, X, **kw)
        F.gp.route = lib.ROUTE_F64
        orac = ocls(F.eq_o, F.gp_o, cast=False)
        orac.u_solve(*args, X, **kw)
        assert np.max(np.abs(prod.last_raw_u - orac.last_raw_u)) < 2e-6
        l2_p, l1_p = _rel_l2(prod.last_raw_u, exact)
        l2_o, l1_o = _rel_l2(orac.last_raw_u, exact)
        assert abs(l2_p - l2_o) <= 1e-6 * l2_o, (l2_p, l2_o)
        assert abs(l1_p - l1_o) <= 1e-6 * l1_o, (l1_p, l1_o)
        assert prod.evaluation_counter == orac.evaluation_counter


def test_fused_sampler_is_bit_identical_to_the_sampler_kernels(mid):
    """On the tcgen05 route the evaluation kernel's loader warps draw the Brownian increments themselves (fused sampler).
    Same Philox addressing, same operation order: switching the fusion off (stand-alone sampler kernels) must not change a bit."""
    F, P = mid, mid.P
    lib = P["lib"]
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(150, 31)                              # 181 points: partial tiles, several CTAs
    X[::7, -1] = F.eq_o.T                                   # degenerate steps included
    for cls, args in ((P["ScaSML"], (3, 3, X)), (P["ScaSMLfh"], (3, None, X, 3))):
        outs = []
        for fused in (True, False):
            s = cls(F.eq, F.gp)
            s.route = lib.ROUTE_TC
            s.quadrature = "gauss_legendre"
            s.fused_sampler = fused
            s.uz_solve(*args)
            outs.append(s.last_raw.copy())
        assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_k_streamed_kernel_in_the_solver(huge):
    """d = 300: the solver's sampled points go through the K-streamed tcgen05 kernel (point images streamed per 64-wide K block)."""
    F, P = huge, huge.P
    lib = P["lib"]
    F.gp.set_right_vector(F.gp_o.right_vector)
    X = F.test_points(20, 5)
    exact = F.eq_o.exact_solution(X)
    prod = P["ScaSML"](F.eq, F.gp)
    assert prod.route is None                                # default: tcgen05 wherever the GP supports it
    prod.quadrature = "gauss_legendre"
    prod.u_solve(2, 2, X)
    orac = ScaSMLOracle(F.eq_o, F.gp_o, cast=False, true_gl=True)
    orac.u_solve(2, 2, X)
    assert np.max(np.abs(prod.last_raw_u - orac.last_raw_u)) < 5e-6
    l2_p, l1_p = _rel_l2(prod.last_raw_u, exact)
    l2_o, l1_o = _rel_l2(orac.last_raw_u, exact)
    assert abs(l2_p - l2_o) <= 1e-5 * l2_o and abs(l1_p - l1_o) <= 1e-5 * l1_o


def test_tc_route_rejects_large_d():
    P = _product()
    lib = P["lib"]
    eq = P["Eq"](1031)
    gp = P["GP"](eq)
    rng = np.random.default_rng(0)
    dom = rng.uniform(-0.5, 0.5, (64, 1031)).astype(np.float16)
    bdy = rng.uniform(-0.5, 0.5, (16, 1031)).astype(np.float16)
    gp._bind(dom, bdy)
    gp.set_right_vector(np.zeros(4 * 64 + 16))
    assert lib.load().scasml_gp_tc_supported(gp._handle) == 0
    gp.route = lib.ROUTE_TC
    with pytest.raises(lib.ScasmlError):
        gp.predict(dom[:4])
    gp.route = lib.ROUTE_F64
    assert gp.predict(dom[:4]).shape == (4, 1)
