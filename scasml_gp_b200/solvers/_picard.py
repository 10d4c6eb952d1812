"""Shared host logic of the four multilevel-Picard solver classes.

The recursion itself (reference solvers/ScaSML.py:149-305 and siblings) runs inside
``scasml_uz_solve``: the tree is enumerated on the host by the library, flattened into level-wise
device batches, and only ``[B, 1+d]`` results come back.  This module holds what stays on the host:
the parameter tables (solvers/ScaSML.py:65-147, computed once per rho instead of once per recursive
call), the bookkeeping the reference keeps in Python ints (``evaluation_counter``, the running
``random.split`` count), batching over test points, and the optional multi-GPU sample sharding
(one ``all_reduce`` of the weighted partial sums).
"""
import ctypes as C
import os

import numpy as np
from scipy.special import lambertw

from .. import _lib


# ---------------------------------------------------------------- tables (solvers/ScaSML.py:65-147) -------
def inverse_gamma(gamma_input):
    c = 0.036534
    L = np.log((gamma_input + c) / np.sqrt(2 * np.pi))
    return np.real(L / np.real(lambertw(L / np.e)) + 0.5)


def lgwt(N, a, b):
    """Legendre-Gauss nodes/weights exactly as the reference computes them -- including its
    ``L[:,1] = y[0,0]`` defect (solvers/ScaSML.py:107): for N >= 2 these are NOT Gauss-Legendre tables,
    N = 2 has a NaN weight and N = 5 does not converge.  Results depend on it, so it is kept."""
    N -= 1
    N1, N2 = N + 1, N + 2
    xu = np.linspace(-1, 1, N1).reshape(1, -1)
    y = np.cos((2 * np.arange(0, N + 1, 1) + 1) * np.pi / (2 * N + 2)) + (0.27 / N1) * np.sin(np.pi * xu * N / N2)
    L = np.zeros((N1, N2))
    y0 = 2
    eps = 2.2204e-16
    iteration, max_iter = 0, 100
    Lp = np.zeros((1, N1))
    with np.errstate(all="ignore"):
        while np.max(np.abs(y - y0)) > eps and iteration < max_iter:
            L[:, 0] = 1
            L[:, 1] = y[0, 0]
            for k in range(2, N1 + 1):
                L[:, k] = (((2 * k - 1) * y * L[:, k - 1] - (k - 1) * L[:, k - 2]) / k)[0]
            Lp = (N2) * (L[:, N1 - 1] - y * L[:, N2 - 1]) / (1 - y * y)
            y0 = y
            y = y0 - L[:, N2 - 1] / Lp
            iteration += 1
        x = (a * (1 - y) + b * (1 + y)) / 2
        w = (b - a) / ((1 - y * y) * (Lp * Lp)) * (N2 * N2) / (N1 * N1)
    return x[0], w[0]


def lgwt_gauss_legendre(N, a, b):
    """True Gauss-Legendre table (flagged deviation; needed where the reference's table is NaN: rho = 1, rho >= 4)."""
    xs, ws = np.polynomial.legendre.leggauss(N)
    xs, ws = xs[::-1], ws[::-1]
    return (a * (1 - xs) + b * (1 + xs)) / 2, ws * (b - a) / 2


def approx_parameters(rhomax, T, quadrature="reference"):
    levels = list(range(1, rhomax + 1))
    Q = np.zeros((rhomax, rhomax), dtype=int)
    Mf = np.zeros((rhomax, rhomax), dtype=int)
    Mg = np.zeros((rhomax, rhomax + 1), dtype=int)
    for rho in range(1, rhomax + 1):
        for k in range(1, levels[rho - 1] + 1):
            Q[rho - 1, k - 1] = int(np.round(inverse_gamma(rho ** (k / 2))))
            Mf[rho - 1, k - 1] = int(np.round(rho ** (k / 2)))
            Mg[rho - 1, k - 1] = int(np.round(rho ** (k - 1)))
        Mg[rho - 1, rho] = rho ** rho
    qmax = int(np.max(Q))
    c = np.zeros((qmax, qmax))
    w = np.zeros((qmax, qmax))
    table = lgwt if quadrature == "reference" else lgwt_gauss_legendre
    for k in range(1, qmax + 1):
        ctemp, wtemp = table(k, 0, T)
        c[:, k - 1] = np.concatenate([ctemp[::-1], np.zeros(qmax - k)])
        w[:, k - 1] = np.concatenate([wtemp[::-1], np.zeros(qmax - k)])
    return Mf, Mg, Q, c, w


# ---------------------------------------------------------------- solver base ------------------------------
class PicardSolverBase(object):
    variant = 0                 # 0 quadrature, 1 full history
    scasml = False              # defect form with the surrogate
    stale_delta = False         # solvers/MLP.py:201,249,270
    route = None                # None: tcgen05 route for the sampled points when the GP supports it (d <= 1022 and float16-valued
                                # collocation points: resident-operand kernel up to d = 126, K-streamed kernel above), else FP64
    quadrature = "reference"    # "reference" (bug-compatible lgwt) or "gauss_legendre" (flagged deviation)
    cast_levels = False         # round inner uz_solve returns to float16 like solvers/ScaSML.py:284
    seed = 0
    workspace_budget_bytes = int(os.environ.get("SCASML_WORKSPACE_BYTES", 24 << 30))
    distributed = False         # shard top-level samples over torch.distributed ranks + one all_reduce
    timing = False              # CUDA-event timing of the sampler / evaluation / reduction groups (last_stats)

    def _init_common(self, equation):
        self.equation = equation
        self.sigma = equation.sigma
        self.mu = equation.mu
        equation.geometry()
        self.T = equation.T
        self.t0 = equation.t0
        self.n_input = equation.n_input
        self.n_output = equation.n_output
        self.evaluation_counter = 0
        self.key = 0                     # running random.split count (reference: self.key = PRNGKey(0), :27)
        self._tables = {}
        self.last_stats = None
        self._last_raw = None
        self._last_raw_dev = None

    # reference-named table helpers (solvers/ScaSML.py:65-147)
    def inverse_gamma(self, gamma_input):
        return inverse_gamma(gamma_input)

    def lgwt(self, N, a, b):
        return lgwt(N, a, b) if self.quadrature == "reference" else lgwt_gauss_legendre(N, a, b)

    def approx_parameters(self, rhomax):
        key = (rhomax, self.quadrature)
        if key not in self._tables:
            self._tables[key] = approx_parameters(rhomax, self.T, self.quadrature)
        return self._tables[key]

    def _clip(self):
        return self.equation.uncertainty if self.scasml else self.equation.norm_estimation

    def _params(self, n, rho, M, rank, world):
        p = _lib.PicardParams()
        p.variant, p.scasml, p.n, p.d = self.variant, int(self.scasml), int(n), self.n_input - 1
        p.M = int(M) if M is not None else 0
        if self.variant == 0:
            if n > rho:
                raise ValueError("quadrature solvers need n <= rho (tables are indexed [rho-1, n])")
            Mf, Mg, Q, c, w = self.approx_parameters(rho)
            self.Mf, self.Mg, self.Q, self.c, self.w = Mf, Mg, Q, c, w       # like solvers/ScaSML.py:161
            qmax = c.shape[0]
            if n > _lib.MAX_LEVEL or qmax > _lib.MAX_Q:
                raise ValueError("level / quadrature size beyond the library limits")
            p.qmax = qmax
            for i in range(n):
                p.Qrow[i] = int(Q[rho - 1, i])
                p.Mfrow[i] = int(Mf[rho - 1, i])
            for i in range(n + 1):
                p.Mgrow[i] = int(Mg[rho - 1, i])
            for k in range(qmax):
                for q in range(qmax):
                    p.c[k * qmax + q] = float(c[k, q])
                    p.w[k * qmax + q] = float(w[k, q])
        else:
            p.qmax = 1
        p.T, p.mu, p.sigma, p.clip = float(self.T), float(self.mu()), float(self.sigma()), float(self._clip())
        p.stale_delta, p.cast_levels = int(self.stale_delta), int(self.cast_levels)
        p.seed, p.key_counter = int(self.seed), int(self.key) & 0xFFFFFFFF
        p.rank, p.world, p.gid0 = int(rank), int(world), 0
        p.timing = int(self.timing)
        p.reserved = 0
        return p

    @property
    def last_raw(self):
        """Un-rounded (u, z) [B, 1 + d] of the last solve (diagnostics / tests); fetched from the device on first access."""
        if self._last_raw is None and self._last_raw_dev is not None:
            self._last_raw = _lib.to_host(self._last_raw_dev)
            self._last_raw_dev = None
        return self._last_raw

    @last_raw.setter
    def last_raw(self, value):
        self._last_raw, self._last_raw_dev = value, None

    def plan(self, n, rho, B, M=None, rank=0, world=1):
        """Host-only tree enumeration: (workspace bytes, stats) for a batch of B rows."""
        p = self._params(n, rho, M, rank, world)
        ws = C.c_size_t(0)
        st = _lib.PicardStats()
        _lib.check(_lib.load().scasml_picard_plan(C.byref(p), int(B), C.byref(ws), C.byref(st)))
        return ws.value, st

    def _dist(self):
        comm = getattr(self, "comm", None)
        if comm is not None:                                  # _lib.AbiComm: the all-reduce through the C ABI
            return comm.rank, comm.world, comm
        if not self.distributed:
            return 0, 1, None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 0, 1, None
        return dist.get_rank(), dist.get_world_size(), dist

    def _uz_device(self, n, rho, x_dev, M=None, with_u_hat=False):
        """x_dev: [B, d+1] float64 CUDA tensor -> [B, 1+d] float64 CUDA tensor (clipped, un-rounded).
        with_u_hat: also return the top-level u_hat(x) [B] (GP.route); in sharded runs every rank evaluates only the rows
        r = rank (mod world) and the values travel in the same all-reduce as the partial sums."""
        lib = _lib.load()
        torch = _lib.torch_cuda()
        _lib.ensure_normal_table()
        B, D = x_dev.shape
        rank, world, dist = self._dist()
        gp_handle = self.GP._handle if self.scasml else C.c_void_p(0)
        if self.scasml:
            self.GP._require_fit()
        route = self.route
        if route is None:     # the top-level u_hat(x) of u_solve always goes through GP.predict (GP.route, FP64 by default)
            route = _lib.ROUTE_TC if (self.scasml and lib.scasml_gp_tc_supported(gp_handle) == 1) else _lib.ROUTE_F64
        buf = torch.empty(B * D + (B if with_u_hat else 0), dtype=torch.float64, device="cuda")
        out = buf[:B * D].view(B, D)
        uh = buf[B * D:] if with_u_hat else None
        p = self._params(n, rho, M, rank, world)
        # batch over test points so the level buffers fit the workspace budget
        # (sharded runs too: every rank sees the same B and budget, so all ranks cut the same chunks; a chunk is a complete
        # sharded solve of its rows -- unit ownership is defined inside the chunk, the RNG is addressed by global row ids --
        # and the partial sums of all chunks go through ONE all-reduce at the end)
        ws1, st1 = self.plan(n, rho, 1, M, 0, 1)
        per_row = max(ws1 // max(world, 1), 1)
        chunk = int(max(1, min(B, self.workspace_budget_bytes // per_row)))
        stats = _lib.PicardStats()
        agg = None
        ws_t = None
        for b0 in range(0, max(B, 1), max(chunk, 1)):
            b1 = min(B, b0 + chunk)
            if b1 <= b0:
                break
            p.gid0 = b0
            need = C.c_size_t(0)
            _lib.check(lib.scasml_picard_plan(C.byref(p), b1 - b0, C.byref(need), None))
            if ws_t is None or ws_t.numel() < need.value:
                ws_t = None
                ws_t = torch.empty(need.value, dtype=torch.uint8, device="cuda")
            _lib.check(lib.scasml_uz_solve(gp_handle, C.byref(p), int(route), _lib.ptr(x_dev[b0:b1]), b1 - b0,
                                           _lib.ptr(out[b0:b1]), _lib.ptr(ws_t), ws_t.numel(), C.byref(stats),
                                           _lib.stream_ptr()))
            if agg is None:
                agg = {f: getattr(stats, f) for f, _ in _lib.PicardStats._fields_}
            else:
                for f in ("executed_points", "launches", "eval_points_total", "eval_launches", "eval_time_ns",
                          "sample_time_ns", "reduce_time_ns", "eval_flops"):
                    agg[f] += getattr(stats, f)
        if B == 0:
            _, st = self.plan(n, rho, 0, M)
            agg = {f: getattr(st, f) for f, _ in _lib.PicardStats._fields_}
        if with_u_hat and B > 0:
            if world > 1:
                uh.zero_()
                mine = torch.arange(rank, B, world, device="cuda")
                if mine.numel():
                    uh[mine] = self.GP._eval(x_dev[mine].contiguous(), _lib.EVAL_U)[0]
            else:
                uh.copy_(self.GP._eval(x_dev, _lib.EVAL_U)[0])
        if world > 1 and (n > 0 or with_u_hat):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)           # the single collective of the path
            if n > 0:
                _lib.check(lib.scasml_clip(_lib.ptr(out), out.numel(), float(self._clip()), _lib.stream_ptr()))
        self.evaluation_counter += int(agg["eval_counter"])
        self.key += int(agg["keys_used"])
        self.last_stats = agg
        return (out, uh) if with_u_hat else out

    def _uz(self, n, rho, x_t, M=None):
        x_dev = _lib.to_device(x_t)
        out = self._uz_device(n, rho, x_dev, M)
        raw = _lib.to_host(out)
        self.last_raw = raw
        return raw

    def _u_solve(self, n, rho, x_t, M=None):
        rank, world, dist = self._dist()
        # one host -> device copy serves the correction and u_hat; sharded runs stage a row slice per rank and all-gather it
        x_dev = _lib.to_device_sharded(x_t, rank, world, dist) if (world > 1 and hasattr(dist, "all_gather_into_tensor")) else _lib.to_device(x_t)
        if self.scasml:
            out, uh = self._uz_device(n, rho, x_dev, M, with_u_hat=True)   # top-level u_hat(x): GP.route (FP64 by default)
        else:
            out = self._uz_device(n, rho, x_dev, M)
        if not self.scasml:
            raw = _lib.to_host(out)
            self.last_raw = raw
            self.last_raw_u = raw[:, :1]
            return raw[:, 0][:, np.newaxis].astype(np.float16)
        torch = _lib.torch_cuda()
        # one device -> host copy (pinned staging) of what u_solve returns: the two columns (u_breve, u_hat).  The un-rounded (u, z) block
        # [B, 1 + d] of the correction stays on the device and comes to the host only if `last_raw` is read (it was 98 % of the D2H bytes).
        both = _lib.to_host(torch.stack((out[:, 0], uh), dim=1))
        u_breve_raw, u_hat_raw = both[:, :1], both[:, 1:]
        self._last_raw, self._last_raw_dev = None, out
        u_breve = u_breve_raw.astype(np.float16)
        self.last_raw_u = u_hat_raw + u_breve_raw
        return u_hat_raw.astype(np.float16) + u_breve             # solvers/ScaSML.py:300-305 (float16 + float16)
