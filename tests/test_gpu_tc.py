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


@pytest.fixture(scope="module")
def edge8():
    return Fitted(d=120, nd=100, nb=30)      # d + 2 = 122: the 8-k-step instantiation (staging rows of 256 bytes)


@pytest.mark.parametrize("which", ["mid", "wide", "edge8", "huge"])
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


def test_guard_bands_around_every_buffer_stay_intact(mid, wide):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer.md), so out-of-bounds writes are hunted with guard bands:
    inputs, outputs and the Picard workspace are carved out of larger allocations whose borders hold a bit pattern that must
    survive every evaluation class on both kernels' shapes (partial point tiles, several tiles per CTA) and a whole solve."""
    import ctypes as C
    for F in (mid, wide):
        P = F.P
        lib, torch = P["lib"], P["torch"]
        gp = F.gp
        gp.set_right_vector(F.gp_o.right_vector)
        G = 4096                                                # guard doubles on either side
        PAT = float(np.frombuffer(np.uint64(0x7FF8DEADBEEF1234).tobytes(), dtype=np.float64)[0])

        def guarded(n):
            buf = torch.full((n + 2 * G,), float("nan"), dtype=torch.float64, device="cuda")
            buf.view(torch.int64).fill_(0x7FF8DEADBEEF1234)
            return buf, buf[G:G + n]

        def intact(buf, n):
            bits = buf.view(torch.int64)
            return bool((bits[:G] == 0x7FF8DEADBEEF1234).all()) and bool((bits[G + n:] == 0x7FF8DEADBEEF1234).all())

        for R in (1, 127, 128 * 150 + 77):                      # 150 tiles on 148 SMs: some CTAs take two tiles
            X = F.test_points(R - R // 6, R // 6) if R > 6 else F.test_points(R, 0)
            R = len(X)
            xbuf, xd = guarded(R * (F.d + 1))
            xd.copy_(torch.from_numpy(X.reshape(-1)).cuda())
            for route in (lib.ROUTE_TC, lib.ROUTE_F64):
                for mode, nout in ((lib.EVAL_U, 1), (lib.EVAL_TERMINAL, 1), (lib.EVAL_UG, 2), (lib.EVAL_PDE, 4)):
                    outs = [guarded(R) for _ in range(nout)]
                    ptrs = [lib.ptr(o[1]) for o in outs] + [C.c_void_p(0)] * (4 - nout)
                    lib.check(lib.load().scasml_gp_eval(gp._handle, lib.ptr(xd), R, mode, route, *ptrs, lib.stream_ptr()))
                    torch.cuda.synchronize()
                    assert all(intact(b, R) for b, _ in outs), (F.d, R, route, mode)
                    assert all(bool(torch.isfinite(v).all()) for _, v in outs), (F.d, R, route, mode)
            assert intact(xbuf, R * (F.d + 1))
        # a whole solve inside a guarded workspace / output
        X = F.test_points(23, 6)
        B, D = X.shape
        xbuf, xd = guarded(B * D)
        xd.copy_(torch.from_numpy(X.reshape(-1)).cuda())
        for cls, n, rho, M in ((P["ScaSML"], 3, 3, None), (P["ScaSMLfh"], 3, None, 3)):
            for route in (lib.ROUTE_TC, lib.ROUTE_F64):
                s = cls(F.eq, gp)
                s.quadrature = "gauss_legendre"
                p = s._params(n, rho, M, 0, 1)
                need = C.c_size_t(0)
                lib.check(lib.load().scasml_picard_plan(C.byref(p), B, C.byref(need), None))
                nws = (need.value + 7) // 8
                wbuf, ws = guarded(nws)
                obuf, out = guarded(B * D)
                st = lib.PicardStats()
                lib.ensure_normal_table()
                lib.check(lib.load().scasml_uz_solve(gp._handle, C.byref(p), route, lib.ptr(xd), B, lib.ptr(out), lib.ptr(ws), need.value,
                                                     C.byref(st), lib.stream_ptr()))
                torch.cuda.synchronize()
                assert intact(wbuf, nws) and intact(obuf, B * D) and intact(xbuf, B * D), (F.d, cls.__name__, route)
                assert bool(torch.isfinite(out).all())


@pytest.mark.parametrize("d,nd,nb", [(20, 20, 6), (20, 90, 30), (100, 40, 8), (100, 300, 60)])
def test_pipeline_protocol_under_many_tiles(d, nd, nb):
    """The resident-operand kernel's mbarrier protocol under load: every CTA runs a dozen point tiles (plus a ragged last one), with one
    pair per class (26 centres: the tile boundary is then as long as the main loop), odd pair counts, and 360 centres.  The kernel is
    deterministic, so repeated launches must agree bit for bit (a parity wait that returns early converts a stale slot; a late one hangs
    until the watchdog traps), and both must agree with the FP64 route."""
    F = Fitted(d=d, nd=nd, nb=nb)
    P, lib, torch = F.P, F.P["lib"], F.P["torch"]
    gp = F.gp
    gp.set_right_vector(F.gp_o.right_vector)
    R = 148 * 128 * 12 + 77
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.uniform(-0.5, 0.5, (R, d)), rng.uniform(0.0, 0.5, (R, 1))], axis=1)
    xd = lib.to_device(X)
    nout = {lib.EVAL_U: 1, lib.EVAL_UG: 2, lib.EVAL_PDE: 4}
    for mode in (lib.EVAL_U, lib.EVAL_UG, lib.EVAL_PDE):
        gp.route = lib.ROUTE_TC
        first = [o.clone() for o in gp._eval(xd, mode, nout=nout[mode])]
        for _ in range(3):
            again = gp._eval(xd, mode, nout=nout[mode])
            torch.cuda.synchronize()
            for a, b in zip(first, again):
                assert not bool(torch.isnan(b).any())
                assert bool((a == b).all()), (d, nd, mode, int((a != b).sum()))
        gp.route = lib.ROUTE_F64
        ref = gp._eval(xd[:4096], mode, nout=nout[mode])
        for a, b in zip(first, ref):
            scale = float(b.abs().max()) + 1e-300
            assert float((a[:4096] - b).abs().max()) < 2e-5 * scale, (d, nd, mode)
    gp.route = lib.ROUTE_F64
