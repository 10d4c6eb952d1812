"""Grad_Dependent_Nonlinear -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates ``equations/equations.py:232-342`` (generator f, terminal g, exact
solution, mu, sigma) and the DeepXDE sampling the reference's ``generate_data``
relies on (``equations/equations.py:344-417``; DeepXDE ``Hypercube`` x
``TimeDomain``: interior points uniform in the cube, boundary points uniform
with one random spatial coordinate snapped to a face, times uniform and
independently permuted; all rounded to float16 like ``dde.config.real``).
"""
import numpy as np


def r16(a):
    """Round to float16 and promote back to float64."""
    return np.asarray(a).astype(np.float16).astype(np.float64)


class EquationOracle:
    def __init__(self, n_input, n_output=1, t0=0.0, T=0.5):
        self.n_input = n_input
        self.n_output = n_output
        self.d = n_input - 1
        self.uncertainty = 1e-1          # equations.py:245
        self.norm_estimation = 1         # equations.py:246
        self.t0, self.T, self.radius = t0, T, 0.5

    def sigma(self, x_t=0):              # equations.py:278-288
        return 0.25

    def mu(self, x_t=0):                 # equations.py:263-276
        s = self.sigma()
        return -1 / self.d - s ** 2 / 2

    def g(self, x_t, cast=True):         # equations.py:248-261
        with np.errstate(over="ignore"):
            res = 1 - 1 / (1 + np.exp(x_t[:, -1] + np.sum(x_t[:, :self.d], axis=1)))
        res = res[:, None]
        return r16(res) if cast else res

    def f(self, x_t, u, z, cast=True):   # equations.py:290-304
        res = self.sigma() * u * np.sum(z, axis=1, keepdims=True)
        return r16(res) if cast else res

    def exact_solution(self, x_t):       # equations.py:306-323
        return self.g(x_t, cast=True)

    # ---- data (DeepXDE semantics restated; float16-valued) ----
    def _points(self, rng, n, boundary):
        d = self.d
        x = rng.random((n, d))
        if boundary:
            dim = rng.integers(0, d, size=n)
            x[np.arange(n), dim] = np.round(x[np.arange(n), dim])
        x = (2 * self.radius) * x - self.radius
        t = rng.random((n, 1)) * (self.T - self.t0) + self.t0
        t = rng.permutation(t)
        return r16(np.hstack([x, t]))

    def generate_data(self, num_domain=100, num_boundary=20, seed=1234):
        rng = np.random.default_rng(seed)
        return self._points(rng, num_domain, False), self._points(rng, num_boundary, True)

    def generate_test_data(self, num_domain=100, num_boundary=20, seed=42):
        return self.generate_data(num_domain, num_boundary, seed)

    # ---- the product's device-side sampler (scasml_geometry_points), restated: same Philox stream, same flat indices ----
    def _points_philox(self, n, boundary, seed, stream_id):
        from . import rng as orng
        d = self.d
        key = orng.make_key(stream_id, 2, seed)
        u = orng.uniforms(key, 0, n * (d + 2)).reshape(n, d + 2)
        x = u[:, :d].copy()
        if boundary:
            face = np.minimum((u[:, d] * d).astype(np.int64), d - 1)
            x[np.arange(n), face] = np.round(x[np.arange(n), face])          # half to even, like rint on the device
        x = (2 * self.radius) * x + (-self.radius)
        t = u[:, d + 1:d + 2] * (self.T - self.t0) + self.t0
        return r16(np.hstack([x, t]))

    def generate_data_philox(self, num_domain=100, num_boundary=20, seed=1234):
        return self._points_philox(num_domain, False, seed, 0), self._points_philox(num_boundary, True, seed, 1)
