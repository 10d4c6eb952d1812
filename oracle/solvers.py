"""Multilevel-Picard solvers -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Float64 restatements of the four recursions, written as the reference's loops:
  * ``ScaSMLOracle``              <- solvers/ScaSML.py:29-63,149-305
  * ``MLPOracle``                 <- solvers/MLP.py:27-55,141-288 (incl. the stale
                                     ``delta_t`` of :201,249,270)
  * ``ScaSMLFullHistoryOracle``   <- solvers/ScaSML_full_history.py:29-72,75-221
  * ``MLPFullHistoryOracle``      <- solvers/MLP_full_history.py:64-196
Randomness: oracle/rng.py (fixed key for terminal draws and for every
full-history draw; running split counter for quadrature step draws, persistent
across calls like ``self.key`` in solvers/ScaSML.py:27,228).  Row r of a call's
batch is the *global* row index, so results do not depend on batching.

``cast=True`` rounds f, g, predict, compute_gradient, compute_PDE_loss and every
uz_solve return to float16 like the reference; ``cast=False`` keeps float64
intermediates (what the CUDA path implements) -- only the public return of
``u_solve`` / ``uz_solve`` is float16 in both modes.  ``last_raw`` keeps the
un-rounded top-level result for 1e-6 parity checks.

``shard=(rank, world)`` restricts the TOP-LEVEL call to the sample units
``u = rank + world*s`` of every (row, sample) array and returns un-clipped
weighted partial sums (``partial_uz``) whose sum over ranks, clipped, equals the
unsharded result.
"""
import numpy as np

from . import rng as orng
from .equation import r16
from .tables import approx_parameters


def _c16(a, cast):
    return r16(a) if cast else a


class _Base:
    variant = None
    scasml = False

    def __init__(self, equation, GP=None, cast=True, seed=0, true_gl=False):
        self.equation = equation
        self.GP = GP
        self.sigma = equation.sigma
        self.mu = equation.mu
        self.T = equation.T
        self.t0 = equation.t0
        self.n_input = equation.n_input
        self.cast = cast
        self.seed = seed
        self.true_gl = true_gl
        self.evaluation_counter = 0
        self.key_counter = 0          # number of random.split calls so far (solvers/ScaSML.py:27,228)
        self._tables = {}
        self.sample_points = 0        # diagnostic: normal vectors drawn over all rows (SURVEY.md 8d work unit)
        self.sample_points_executed = 0   # same, without the level-0 terminals the reference draws and discards

    # -- generator / terminal (solvers/ScaSML.py:29-63 vs solvers/MLP.py:27-55) --
    def f(self, x_t, u, z):
        eq, cast = self.equation, self.cast
        if not self.scasml:
            return eq.f(x_t, u, z, cast=cast)
        self.evaluation_counter += 1
        u_hat = _c16(self.GP.predict_raw(x_t)[:, None], cast)
        grad = _c16(self.GP.gradient_raw(x_t), cast)[:, :-1]
        val1 = eq.f(x_t, u + u_hat, eq.sigma() * grad + z, cast=cast)
        val2 = eq.f(x_t, u_hat, eq.sigma() * grad, cast=cast)
        return val1 - val2

    def g(self, x_t):
        eq, cast = self.equation, self.cast
        if not self.scasml:
            return eq.g(x_t, cast=cast)[:, 0]
        self.evaluation_counter += 1
        u_hat = _c16(self.GP.predict_raw(x_t)[:, None], cast)
        return (eq.g(x_t, cast=cast) - u_hat)[:, 0]

    def _pde(self, x_t):
        return _c16(self.GP.pde_raw(x_t)[:, None], self.cast)

    def _clip(self):
        return self.equation.uncertainty if self.scasml else self.equation.norm_estimation

    # -- sharding helper: mask [rows, MC] of the units owned at the top level --
    @staticmethod
    def _own(rows, MC, shard):
        if shard is None:
            return None
        rank, world = shard
        u = np.arange(rows * MC).reshape(rows, MC)
        return (u % world) == rank


class _Quadrature(_Base):
    variant = "quadrature"

    def tables(self, rho):
        if rho not in self._tables:
            self._tables[rho] = approx_parameters(rho, self.T, true_gl=self.true_gl)
        return self._tables[rho]

    def uz_solve(self, n, rho, x_t, gid0=0, shard=None, _top=True):
        Mf, Mg, Q, c, w = self.tables(rho)
        eq, T = self.equation, self.T
        x_t = np.asarray(x_t, dtype=np.float64)
        dim = self.n_input - 1
        B = x_t.shape[0]
        sigma, mu = self.sigma(), self.mu()
        x, t = x_t[:, :-1], x_t[:, -1]
        cast = self.cast
        # solvers/ScaSML.py:174-175 (same op order: ((T-t)*c)/T + t)
        cloc = (T - t)[:, None, None] * c[None, :] / T + t[:, None, None]
        wloc = (T - t)[:, None, None] * w[None, :] / T
        MC_g = int(Mg[rho - 1, n])
        keyT = orng.make_key(0, orng.DOMAIN_FIXED, self.seed)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            if n == 0:
                # solvers/ScaSML.py:190-219: the terminal work is done and discarded.
                if self.scasml:
                    self.evaluation_counter += 1
                self.evaluation_counter += MC_g
                self.sample_points += B * MC_g
                return np.zeros((B, dim + 1))
            own = self._own(B, MC_g, shard)
            std_normal = orng.normals(keyT, gid0 * MC_g * dim, B * MC_g * dim).reshape(B, MC_g, dim)
            dW = np.sqrt(T - t)[:, None, None] * std_normal
            X = x[:, None, :] + mu * (T - t)[:, None, None] + sigma * dW
            term_in = np.concatenate([X, np.full((B, MC_g, 1), T)], axis=2).reshape(-1, self.n_input)
            gv = self.g(term_in).reshape(B, MC_g, 1)
            self.evaluation_counter += MC_g
            self.sample_points += B * MC_g
            self.sample_points_executed += B * MC_g
            if own is not None:
                gv = gv * own[:, :, None]
            u = gv.sum(axis=1) / MC_g
            delta_t = (T - t + 1e-6)[:, None]
            z = (gv * std_normal).sum(axis=1) / MC_g / delta_t
            for l in range(n):
                q = int(Q[rho - 1, n - l - 1])
                d = cloc[:, :q, q - 1] - np.concatenate([t[:, None], cloc[:, :q - 1, q - 1]], axis=1)
                MC_f = int(Mf[rho - 1, n - l - 1])
                own = self._own(B, MC_f, shard)
                X = np.repeat(x[:, None, :], MC_f, axis=1)
                W = np.zeros((B, MC_f, dim))
                for k in range(q):
                    key = orng.make_key(self.key_counter, orng.DOMAIN_STEP, self.seed)
                    self.key_counter += 1
                    dWr = orng.normals(key, gid0 * MC_f * dim, B * MC_f * dim).reshape(B, MC_f, dim)
                    dW = np.sqrt(d[:, k])[:, None, None] * dWr
                    W = W + dW
                    X = X + (mu * d[:, k][:, None, None] + sigma * dW)
                    self.sample_points += B * MC_f
                    self.sample_points_executed += B * MC_f
                    tk = cloc[:, k, q - 1]
                    pts = np.concatenate([X, np.repeat(tk[:, None, None], MC_f, axis=1)], axis=2).reshape(-1, self.n_input)
                    sim = self.uz_solve(l, rho, pts, gid0 * MC_f, None, _top=False).reshape(B, MC_f, -1)
                    y = self.f(pts, sim[:, :, 0].reshape(-1, 1), sim[:, :, 1:].reshape(-1, dim)).reshape(B, MC_f, 1)
                    self.evaluation_counter += MC_f
                    dt_add, dt_sub = self._deltas(delta_t, tk, t)
                    if own is not None:
                        y = y * own[:, :, None]
                    wk = wloc[:, k, q - 1][:, None]
                    u = u + wk * (y.sum(axis=1) / MC_f)
                    z = z + wk * (y * W).sum(axis=1) / (MC_f * dt_add)
                    delta_t = dt_add
                    if l:
                        sim = self.uz_solve(l - 1, rho, pts, gid0 * MC_f, None, _top=False).reshape(B, MC_f, -1)
                        y = self.f(pts, sim[:, :, 0].reshape(-1, 1), sim[:, :, 1:].reshape(-1, dim)).reshape(B, MC_f, 1)
                        self.evaluation_counter += MC_f
                        if own is not None:
                            y = y * own[:, :, None]
                        u = u - wk * (y.sum(axis=1) / MC_f)
                        delta_t = dt_sub
                        z = z - wk * (y * W).sum(axis=1) / (MC_f * delta_t)
                    elif self.scasml:
                        eps = self._pde(pts).reshape(B, MC_f, 1)
                        if own is not None:
                            eps = eps * own[:, :, None]
                        u = u + wk * (eps.sum(axis=1) / MC_f)
                        delta_t = dt_sub
                        z = z + wk * (eps * W).sum(axis=1) / (MC_f * delta_t)
            out = np.concatenate([u, z], axis=-1)
            if _top:
                self.partial_uz = out
                if shard is not None:
                    return out
            lim = self._clip()
            out = np.clip(out, -lim, lim)
            if _top:
                self.last_raw = out
                return out.astype(np.float16)
            return _c16(out, cast)

    def u_solve(self, n, rho, x_t):
        uz = self.uz_solve(n, rho, x_t)
        u_breve = uz[:, 0][:, None]
        if not self.scasml:
            self.last_raw_u = self.last_raw[:, :1]
            return u_breve                                            # solvers/MLP.py:276-288
        u_hat_raw = self.GP.predict_raw(np.asarray(x_t, dtype=np.float64))[:, None]
        self.last_raw_u = u_hat_raw + self.last_raw[:, :1]
        return u_hat_raw.astype(np.float16) + u_breve                 # solvers/ScaSML.py:300-305


class ScaSMLOracle(_Quadrature):
    scasml = True

    def _deltas(self, delta_prev, tk, t):
        dt = (tk - t + 1e-6)[:, None]                                 # solvers/ScaSML.py:253,272,279
        return dt, dt


class MLPOracle(_Quadrature):
    scasml = False

    def _deltas(self, delta_prev, tk, t):
        # solvers/MLP.py:249 uses whatever delta_t currently holds; :270 refreshes it in the l>0 branch.
        return delta_prev, (tk - t + 1e-6)[:, None]

    def uz_solve(self, n, rho, x_t, gid0=0, shard=None, _top=True):
        return super().uz_solve(n, rho, x_t, gid0, shard, _top)


class _FullHistory(_Base):
    variant = "full_history"

    def uz_solve(self, n, rho, x_t, M, gid0=0, shard=None, _top=True):
        eq, T = self.equation, self.T
        x_t = np.asarray(x_t, dtype=np.float64)
        dim = self.n_input - 1
        B = x_t.shape[0]
        sigma, mu = self.sigma(), self.mu()
        x, t = x_t[:, :-1], x_t[:, -1]
        cast = self.cast
        keyT = orng.make_key(0, orng.DOMAIN_FIXED, self.seed)
        MC_g = int(M ** n)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            if n == 0:
                if self.scasml:
                    self.evaluation_counter += 1
                self.evaluation_counter += MC_g
                self.sample_points += B * MC_g
                return np.zeros((B, dim + 1))
            own = self._own(B, MC_g, shard)
            std_normal = orng.normals(keyT, gid0 * MC_g * dim, B * MC_g * dim).reshape(B, MC_g, dim)
            dW = np.sqrt(T - t)[:, None, None] * std_normal
            X = x[:, None, :] + mu * (T - t)[:, None, None] + sigma * dW
            term_in = np.concatenate([X, np.full((B, MC_g, 1), T)], axis=2).reshape(-1, self.n_input)
            gv = self.g(term_in).reshape(B, MC_g, 1)
            self.evaluation_counter += MC_g
            self.sample_points += B * MC_g
            self.sample_points_executed += B * MC_g
            if own is not None:
                gv = gv * own[:, :, None]
            u = gv.sum(axis=1) / MC_g
            delta_t = (T - t)[:, None]                               # solvers/ScaSML_full_history.py:133 (no +1e-6)
            z = (gv * std_normal).sum(axis=1) / MC_g / delta_t
            for l in range(n):
                MC_f = int(M ** (n - l))
                own = self._own(B, MC_f, shard)
                tau = orng.uniforms(keyT, gid0 * MC_f, B * MC_f).reshape(B, MC_f)
                steps = (tau * (T - t)[:, None]).reshape(B, MC_f, 1)
                std_normal = orng.normals(keyT, gid0 * MC_f * dim, B * MC_f * dim).reshape(B, MC_f, dim)
                dW = np.sqrt(steps) * std_normal
                X = x[:, None, :] + (mu * steps + sigma * dW)
                self.sample_points += B * MC_f
                self.sample_points_executed += B * MC_f
                pts = np.concatenate([X, t[:, None, None] + steps], axis=2).reshape(-1, self.n_input)
                sim = self.uz_solve(l, None, pts, M, gid0 * MC_f, None, _top=False).reshape(B, MC_f, dim + 1)
                y = self.f(pts, sim[:, :, 0].reshape(-1, 1), sim[:, :, 1:].reshape(-1, dim)).reshape(B, MC_f, 1)
                self.evaluation_counter += MC_g if self.scasml else MC_f   # quirk A.3-7
                if own is not None:
                    y = y * own[:, :, None]
                Tt = (T - t)[:, None]
                dsq = np.sqrt(steps + 1e-6)
                u = u + Tt * (y.sum(axis=1) / MC_f)
                z = z + Tt * ((y * std_normal / dsq).sum(axis=1) / MC_f)
                if l:
                    sim = self.uz_solve(l - 1, None, pts, M, gid0 * MC_f, None, _top=False).reshape(B, MC_f, dim + 1)
                    y = self.f(pts, sim[:, :, 0].reshape(-1, 1), sim[:, :, 1:].reshape(-1, dim)).reshape(B, MC_f, 1)
                    self.evaluation_counter += MC_g if self.scasml else MC_f
                    if own is not None:
                        y = y * own[:, :, None]
                    u = u - Tt * (y.sum(axis=1) / MC_f)
                    z = z - Tt * ((y * std_normal / dsq).sum(axis=1) / MC_f)
                elif self.scasml:
                    eps = self._pde(pts).reshape(B, MC_f, 1)
                    if own is not None:
                        eps = eps * own[:, :, None]
                    u = u + Tt * (eps.sum(axis=1) / MC_f)
                    z = z + Tt * ((eps * std_normal / dsq).sum(axis=1) / MC_f)
            out = np.concatenate([u, z], axis=-1)
            if _top:
                self.partial_uz = out
                if shard is not None:
                    return out
            lim = self._clip()
            out = np.clip(out, -lim, lim)
            if _top:
                self.last_raw = out
                return out.astype(np.float16)
            # ScaSML_full_history returns un-cast (:199); MLP_full_history casts (:179)
            return out if self.scasml else _c16(out, cast)

    def u_solve(self, n, rho, x_t, M=3):
        uz = self.uz_solve(n, rho, x_t, M)
        u_breve = uz[:, 0][:, None]
        if not self.scasml:
            self.last_raw_u = self.last_raw[:, :1]
            return u_breve
        u_hat_raw = self.GP.predict_raw(np.asarray(x_t, dtype=np.float64))[:, None]
        self.last_raw_u = u_hat_raw + self.last_raw[:, :1]
        return u_hat_raw.astype(np.float16) + u_breve


class ScaSMLFullHistoryOracle(_FullHistory):
    scasml = True


class MLPFullHistoryOracle(_FullHistory):
    scasml = False
