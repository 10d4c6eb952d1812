"""PDE-constrained Gaussian-process surrogate -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Float64 restatement of ``models/GP.py``:
  * Gaussian kernel and its 15 derivative functionals (``models/GP.py:41-180``)
    in closed form, including the rotated-coordinate 5-index "Laplacian"
    (``:28-39,91-93,101-103``): SURVEY.md App. B, re-derived and checked against a
    torch-autograd restatement of the reference's nested ``jax.grad`` code in
    ``tests/test_oracle_closed_forms.py``;
  * Gram matrix, 25 blocks (``:182-258``), entries rounded to float16 like ``:258``;
  * the fit (``:430-444,487-604,705-719``): J(sol) = b^T (K+nu I)^{-1} b, damped
    Newton with the closed-form gradient/Hessian, alpha = (K+nu I)^{-1} z;
  * predict / compute_gradient / compute_PDE_loss (``:630-687,746-769``).
Deviations from the reference (declared): kernel-row entries
are not individually rounded to float16; the SVD "Cholesky" factor stored in
float16 (``:260-268``) is replaced by the exact (K + nu I)^{-1}.
"""
import numpy as np
import scipy.linalg as sla

from .equation import r16

MC = 5  # models/GP.py:30


class _Pairs:
    """Pairwise geometry between row points X [R,D] and column points Y [N,D] (time last)."""

    def __init__(self, X, Y, a, idx, d):
        self.X, self.Y, self.a, self.I, self.d = X, Y, a, np.asarray(idx), d
        self._c = {}

    def _get(self, name, fn):
        if name not in self._c:
            self._c[name] = fn()
        return self._c[name]

    @staticmethod
    def _roll(Z):                      # roll(z) = (z_1, ..., z_d, z_0)
        return np.concatenate([Z[:, 1:], Z[:, :1]], axis=1)

    def _sq(self, A, B):
        na = (A * A).sum(1)[:, None]
        nb = (B * B).sum(1)[None, :]
        return na + nb - 2.0 * (A @ B.T)

    # plain pair r = x - y
    @property
    def k(self):
        return self._get("k", lambda: np.exp(-0.5 * self.a * self._sq(self.X, self.Y)))

    @property
    def S(self):
        d = self.d
        return self._get("S", lambda: self.X[:, :d].sum(1)[:, None] - self.Y[:, :d].sum(1)[None, :])

    @property
    def rt(self):
        d = self.d
        return self._get("rt", lambda: self.X[:, d][:, None] - self.Y[:, d][None, :])

    # ry = x - roll(y)
    @property
    def ky(self):
        return self._get("ky", lambda: np.exp(-0.5 * self.a * self._sq(self.X, self._roll(self.Y))))

    @property
    def Sy(self):
        d = self.d
        return self._get("Sy", lambda: self.X[:, :d].sum(1)[:, None] - self.Y[:, 1:d + 1].sum(1)[None, :])

    @property
    def ry_I(self):                    # [R,N,MC]
        return self._get("ryI", lambda: self.X[:, None, self.I] - self.Y[None, :, self.I + 1])

    @property
    def ry_d(self):
        d = self.d
        return self._get("ryd", lambda: self.X[:, d][:, None] - self.Y[:, 0][None, :])

    # rx = roll(x) - y
    @property
    def kx(self):
        return self._get("kx", lambda: np.exp(-0.5 * self.a * self._sq(self._roll(self.X), self.Y)))

    @property
    def Sx(self):
        d = self.d
        return self._get("Sx", lambda: self.X[:, 1:d + 1].sum(1)[:, None] - self.Y[:, :d].sum(1)[None, :])

    @property
    def rx_I(self):
        return self._get("rxI", lambda: self.X[:, None, self.I + 1] - self.Y[None, :, self.I])

    @property
    def rx_d(self):
        d = self.d
        return self._get("rxd", lambda: self.X[:, 0][:, None] - self.Y[:, d][None, :])

    # q = roll(x) - roll(y)
    @property
    def q2(self):                      # sum_{n in I} q_n^2
        def fn():
            q = self.X[:, None, self.I + 1] - self.Y[None, :, self.I + 1]
            return (q * q).sum(2)
        return self._get("q2", fn)

    def block(self, rowop, colop):
        """Functional  rowop_x colop_y kappa(x, y)  for every pair, [R,N]."""
        a, d = self.a, self.d
        key = (rowop, colop)
        if key == ("id", "id"):
            return self.k
        if key == ("id", "dt"):
            return a * self.rt * self.k
        if key == ("id", "div"):
            return a * self.S * self.k
        if key == ("id", "lap"):
            return d * (a * a * (self.ry_I ** 2) - a).mean(2) * self.ky
        if key == ("dt", "id"):
            return -a * self.rt * self.k
        if key == ("div", "id"):
            return -a * self.S * self.k
        if key == ("lap", "id"):
            return d * (a * a * (self.rx_I ** 2) - a).mean(2) * self.kx
        if key == ("dt", "dt"):
            return (a - a * a * self.rt ** 2) * self.k
        if key in (("dt", "div"), ("div", "dt")):
            return -a * a * self.rt * self.S * self.k
        if key == ("div", "div"):
            return (a * d - a * a * self.S ** 2) * self.k
        if key == ("dt", "lap"):
            return -a * self.ry_d * d * (a * a * (self.ry_I ** 2) - a).mean(2) * self.ky
        if key == ("div", "lap"):
            ry, Sy = self.ry_I, self.Sy[:, :, None]
            return d * (2 * a * a * ry + a * a * Sy - a ** 3 * Sy * ry ** 2).mean(2) * self.ky
        if key == ("lap", "dt"):
            return a * self.rx_d * d * (a * a * (self.rx_I ** 2) - a).mean(2) * self.kx
        if key == ("lap", "div"):
            rx, Sx = self.rx_I, self.Sx[:, :, None]
            return d * (-2 * a * a * rx - a * a * Sx + a ** 3 * Sx * rx ** 2).mean(2) * self.kx
        if key == ("lap", "lap"):
            q2 = self.q2
            A = a * a * q2 - MC * a
            return (d * d / (MC * MC)) * self.k * (A * A + 2 * MC * a * a - 4 * a ** 3 * q2)
        raise KeyError(key)


COLOPS = ("id", "id", "lap", "dt", "div")     # column order [D | B | lap(D) | dt(D) | div(D)], models/GP.py:251-258
ROWOPS = ("id", "id", "lap", "dt", "div")


class GPOracle:
    """Restatement of ``GP_Grad_Dependent_Nonlinear`` (models/GP.py:693-769)."""

    def __init__(self, equation, idx_set=None, nugget=1e-2, cast=True):
        self.equation = equation
        self.d = equation.d
        self.n_input = equation.n_input
        self.sigma = equation.sigma() * np.sqrt(self.d)          # models/GP.py:25
        self.a = 1.0 / (self.sigma ** 2)
        self.nugget = nugget                                     # models/GP.py:26
        if idx_set is None:
            idx_set = np.random.default_rng(0).choice(self.d, MC, replace=False)
        self.idx_set = np.asarray(idx_set, dtype=np.int64)
        self.cast = cast
        self.loss_history = []

    # ---- Gram (models/GP.py:182-268) ----
    def _sets(self):
        return [self.x_t_domain, self.x_t_boundary, self.x_t_domain, self.x_t_domain, self.x_t_domain]

    def gram(self, x_dom, x_bdy, f16_entries=True):
        self.x_t_domain = np.asarray(x_dom, dtype=np.float64)
        self.x_t_boundary = np.asarray(x_bdy, dtype=np.float64)
        self.N_domain, self.N_boundary = len(x_dom), len(x_bdy)
        self.phi_dim = 4 * self.N_domain + self.N_boundary
        cache = {}

        def pairs(A, B, ia, ib):
            if (ia, ib) not in cache:
                cache[(ia, ib)] = _Pairs(A, B, self.a, self.idx_set, self.d)
            return cache[(ia, ib)]

        sets = self._sets()
        which = [0, 1, 0, 0, 0]
        rows = []
        for i in range(5):
            row = []
            for j in range(5):
                P = pairs(sets[i], sets[j], which[i], which[j])
                row.append(P.block(ROWOPS[i], COLOPS[j]))
            rows.append(np.hstack(row))
        K = np.vstack(rows)
        return r16(K) if f16_entries else K

    # ---- fit (models/GP.py:430-444, 487-604, 705-719) ----
    def time_der_rep(self, sol):
        N = self.N_domain
        s2 = self.equation.sigma() ** 2
        z1, z3, z5 = sol[:N], sol[N:2 * N], sol[2 * N:]
        return -s2 * z1 * z5 + (1 / self.d + s2 / 2) * z5 - (s2 / 2) * z3

    def _b(self, sol):
        N = self.N_domain
        return np.concatenate([sol[:N], self.g_bdy, sol[N:2 * N], self.time_der_rep(sol), sol[2 * N:]])

    def GPsolver(self, x_dom, x_bdy, GN_steps=20, sol0=None, damping=1e-4, tol=1e-5):
        K = self.gram(x_dom, x_bdy)
        N, Nb = self.N_domain, self.N_boundary
        self.g_bdy = self.equation.g(self.x_t_boundary, cast=True)[:, 0]       # models/GP.py:417-419 (eq.g is f16)
        Kp = K + self.nugget * np.eye(self.phi_dim)
        cho = sla.cho_factor(Kp, lower=True)
        P = sla.cho_solve(cho, np.eye(self.phi_dim))
        P = 0.5 * (P + P.T)
        if sol0 is None:
            sol0 = np.random.default_rng(0).standard_normal(3 * N) * 1e-3     # models/GP.py:501 (stream differs)
        sol = np.asarray(sol0, dtype=np.float64).copy()
        s2 = self.equation.sigma() ** 2
        c1 = 1 / self.d + s2 / 2
        o1, o3, o4, o5 = 0, N + Nb, 2 * N + Nb, 3 * N + Nb       # block offsets in phi: z1, z3, F, z5
        offs = [o1, o3, o5]

        def Pblk(r, c):
            return P[r:r + N, c:c + N]

        hist = []
        b = self._b(sol)
        hist.append(float(b @ (P @ b)))
        self.newton_steps = 0
        for _ in range(GN_steps):
            z1, z5 = sol[:N], sol[2 * N:]
            b = self._b(sol)
            w = P @ b
            wF = w[o4:o4 + N]
            D = [-s2 * z5, -(s2 / 2) * np.ones(N), -s2 * z1 + c1]     # dF/dz1, dF/dz3, dF/dz5 (diagonals)
            grad = 2 * np.concatenate([w[offs[i]:offs[i] + N] + D[i] * wF for i in range(3)])
            if np.linalg.norm(grad) < tol:                            # models/GP.py:521
                break
            H = np.empty((3 * N, 3 * N))
            P44 = Pblk(o4, o4)
            for i in range(3):
                for j in range(3):
                    blk = Pblk(offs[i], offs[j]) + D[i][:, None] * Pblk(o4, offs[j]) \
                        + Pblk(offs[i], o4) * D[j][None, :] + D[i][:, None] * P44 * D[j][None, :]
                    H[i * N:(i + 1) * N, j * N:(j + 1) * N] = 2 * blk
            idx = np.arange(N)
            H[idx, 2 * N + idx] += 2 * (-s2) * wF
            H[2 * N + idx, idx] += 2 * (-s2) * wF
            H[np.arange(3 * N), np.arange(3 * N)] += damping        # models/GP.py:529
            sol = sol + np.linalg.solve(H, -grad)                     # models/GP.py:533,573
            b = self._b(sol)
            hist.append(float(b @ (P @ b)))
            self.newton_steps += 1
        self.loss_history = hist
        self.sol = sol
        z = self._b(sol)                                              # models/GP.py:593-598
        self.right_vector = (P @ z)[:, None]                          # models/GP.py:599-600
        self._P = P
        return self.predict(self.x_t_domain)

    # ---- inference (models/GP.py:630-687, 746-769) ----
    def _alpha_blocks(self):
        N, Nb = self.N_domain, self.N_boundary
        al = self.right_vector[:, 0]
        return al[:N], al[N:N + Nb], al[N + Nb:2 * N + Nb], al[2 * N + Nb:3 * N + Nb], al[3 * N + Nb:]

    def _row_dot(self, X, rowop):
        """(rowop_x applied to the kernel row at x) @ alpha, all 5 column blocks."""
        a1, a2, a3, a4, a5 = self._alpha_blocks()
        PD = _Pairs(X, self.x_t_domain, self.a, self.idx_set, self.d)
        PB = _Pairs(X, self.x_t_boundary, self.a, self.idx_set, self.d)
        return (PD.block(rowop, "id") @ a1 + PB.block(rowop, "id") @ a2 + PD.block(rowop, "lap") @ a3
                + PD.block(rowop, "dt") @ a4 + PD.block(rowop, "div") @ a5)

    def _chunked(self, fn, X, chunk=4096):
        X = np.asarray(X, dtype=np.float64)
        if len(X) <= chunk:
            return fn(X)
        return np.concatenate([fn(X[i:i + chunk]) for i in range(0, len(X), chunk)], axis=0)

    def predict_raw(self, X):
        return self._chunked(lambda Z: self._row_dot(Z, "id"), X)

    def gradient_raw(self, X):
        """grad_x u_hat, [R, d+1]."""
        def fn(Z):
            a, d, I = self.a, self.d, self.idx_set
            a1, a2, a3, a4, a5 = self._alpha_blocks()
            YD, YB = self.x_t_domain, self.x_t_boundary
            PD = _Pairs(Z, YD, a, I, d)
            PB = _Pairs(Z, YB, a, I, d)
            cD = PD.k * (a1 + a4 * a * PD.rt + a5 * a * PD.S)          # r-type coefficients, domain
            cB = PB.k * a2
            MH = (a * a * PD.ry_I ** 2 - a).mean(2)
            cL = PD.ky * a3 * d * MH                                   # ry-type coefficients
            g = -a * (Z * (cD.sum(1) + cB.sum(1) + cL.sum(1))[:, None]
                      - cD @ YD - cB @ YB - cL @ _Pairs._roll(YD))
            g[:, d] += a * (PD.k @ a4)                                 # e_t term of grad dt_y
            g[:, :d] += (a * (PD.k @ a5))[:, None]                     # 1_s term of grad div_y
            lapI = np.einsum("rn,rnm->rm", PD.ky * a3 * d * (2 * a * a / MC), PD.ry_I)
            g[:, I] += lapI                                            # 1_I term of grad lap_y (indices distinct)
            return g
        return self._chunked(fn, X)

    def pde_raw(self, X):
        def fn(Z):
            s = self.equation.sigma()
            d = self.d
            u = self._row_dot(Z, "id")
            dv = self._row_dot(Z, "div")
            lp = self._row_dot(Z, "lap")
            dt = self._row_dot(Z, "dt")
            return dt + (s * s * u - 1 / d - s * s / 2) * dv + (s * s / 2) * lp, u, dv, lp, dt
        if len(X) <= 2048:
            return fn(np.asarray(X, dtype=np.float64))[0]
        return np.concatenate([fn(np.asarray(X[i:i + 2048], dtype=np.float64))[0] for i in range(0, len(X), 2048)])

    def pde_terms_raw(self, X):
        """(eps, u, div_x u, lap_x u, dt u) -- for kernel-level parity tests."""
        X = np.asarray(X, dtype=np.float64)
        s, d = self.equation.sigma(), self.d
        u, dv, lp, dt = (self._row_dot(X, op) for op in ("id", "div", "lap", "dt"))
        return dt + (s * s * u - 1 / d - s * s / 2) * dv + (s * s / 2) * lp, u, dv, lp, dt

    # public API (float16 returns, like the reference)
    def predict(self, X):
        return self.predict_raw(X)[:, None].astype(np.float16)

    def compute_gradient(self, X, sol=None):
        return self.gradient_raw(X).astype(np.float16)

    def compute_PDE_loss(self, X):
        return self.pde_raw(X)[:, None].astype(np.float16)
