"""Problem definition -- host mirror of the reference's ``equations/equations.py``.

``Grad_Dependent_Nonlinear`` keeps the reference's constructor, attributes and method names
(equations/equations.py:232-417).  f / g / exact_solution run on the device through the C ABI
(``scasml_equation_f`` / ``scasml_equation_g``) and return float16 NumPy arrays like the reference's
``.astype(jnp.float16)``.  Geometry sampling restates the DeepXDE calls the reference makes
(``Hypercube x TimeDomain``, ``random_points`` / ``random_boundary_points``, float16 ``config.real``) on
NumPy's global generator, so ``np.random.seed(1234)`` in a launcher has the same role as in the reference.
"""
import numpy as np

from .. import _lib


class Equation(object):
    """Base class: only what the hot path uses (equations/equations.py:15-230 is PINN-era scaffolding)."""

    def __init__(self, n_input, n_output=1):
        self.n_input = n_input    # dimension of the input, including time
        self.n_output = n_output  # dimension of the output

    def g(self, x_t):
        """Terminal constraint dispatch (equations/equations.py:146-162)."""
        if hasattr(self, 'terminal_constraint'):
            return self.terminal_constraint(x_t)
        raise NotImplementedError

    def f(self, x_t, u, z):
        raise NotImplementedError

    def mu(self, x_t=0):
        raise NotImplementedError

    def sigma(self, x_t=0):
        raise NotImplementedError


class _Hypercube(object):
    """DeepXDE ``geometry.Hypercube`` restricted to the two samplers the reference calls."""

    def __init__(self, xmin, xmax):
        self.xmin = np.asarray(xmin, dtype=np.float64)
        self.xmax = np.asarray(xmax, dtype=np.float64)
        self.dim = len(self.xmin)

    def random_points(self, n):
        x = np.random.random(size=(n, self.dim))
        return (self.xmax - self.xmin) * x + self.xmin

    def random_boundary_points(self, n):
        x = np.random.random(size=(n, self.dim))
        rand_dim = np.random.randint(self.dim, size=n)
        x[np.arange(n), rand_dim] = np.round(x[np.arange(n), rand_dim])
        return (self.xmax - self.xmin) * x + self.xmin


class _TimeDomain(object):
    def __init__(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def random_points(self, n):
        return np.random.random(size=(n, 1)) * (self.t1 - self.t0) + self.t0


class _GeometryXTime(object):
    def __init__(self, geometry, timedomain):
        self.geometry, self.timedomain = geometry, timedomain

    def _with_time(self, x):
        t = np.random.permutation(self.timedomain.random_points(len(x)))
        return np.hstack((x, t)).astype(np.float16)          # dde.config.set_default_float("float16")

    def random_points(self, n):
        return self._with_time(self.geometry.random_points(n))

    def random_boundary_points(self, n):
        return self._with_time(self.geometry.random_boundary_points(n))


class Grad_Dependent_Nonlinear(Equation):
    '''High-dimensional semilinear PDE with exact solution sigmoid(t + sum x) (equations/equations.py:232).'''

    def __init__(self, n_input, n_output=1):
        super().__init__(n_input, n_output)
        self.uncertainty = 1e-1       # equations.py:245
        self.norm_estimation = 1      # equations.py:246

    # ---- device-evaluated closed forms ----
    def _g_device(self, x_t):
        lib = _lib.load()
        x = _lib.to_device(x_t)
        torch = _lib.torch_cuda()
        out = torch.empty(x.shape[0], dtype=torch.float64, device="cuda")
        _lib.check(lib.scasml_equation_g(_lib.ptr(x), x.shape[0], self.n_input - 1, _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()

    def terminal_constraint(self, x_t):
        '''1 - 1/(1 + exp(t + sum_i x_i)), shape (batch, 1), float16 (equations.py:248-261).'''
        return self._g_device(x_t)[:, np.newaxis].astype(np.float16)

    def mu(self, x_t=0):
        '''Drift (equations.py:263-276).'''
        sigma = self.sigma()
        d = self.n_input - 1
        return -1 / d - sigma ** 2 / 2

    def sigma(self, x_t=0):
        '''Diffusion (equations.py:278-288).'''
        return 0.25

    def f(self, x_t, u, z):
        '''Generator sigma * u * sum_i z_i, shape (batch, 1), float16 (equations.py:290-304).'''
        lib = _lib.load()
        torch = _lib.torch_cuda()
        ud = _lib.to_device(np.asarray(u, dtype=np.float64).reshape(-1))
        zd = _lib.to_device(np.asarray(z, dtype=np.float64).reshape(ud.shape[0], -1))
        out = torch.empty(ud.shape[0], dtype=torch.float64, device="cuda")
        _lib.check(lib.scasml_equation_f(_lib.ptr(ud), _lib.ptr(zd), ud.shape[0], zd.shape[1], self.sigma(),
                                         _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()[:, np.newaxis].astype(np.float16)

    def exact_solution(self, x_t):
        '''Exact solution (equations.py:306-323): same closed form as the terminal constraint.'''
        return self._g_device(x_t)[:, np.newaxis].astype(np.float16)

    # ---- geometry / data (equations.py:344-417) ----
    def geometry(self, t0=0, T=0.5):
        self.t0 = t0
        self.T = T
        self.radius = 0.5
        spacedomain = _Hypercube([-self.radius] * (self.n_input - 1), [self.radius] * (self.n_input - 1))
        timedomain = _TimeDomain(t0, self.T)
        self.geomx = spacedomain
        self.geomt = timedomain
        return _GeometryXTime(spacedomain, timedomain)

    def test_geometry(self, t0=0, T=0.5):
        self.t0 = t0
        self.T = T
        self.test_T = T
        self.test_radius = 0.5
        spacedomain = _Hypercube([-self.test_radius] * (self.n_input - 1), [self.test_radius] * (self.n_input - 1))
        timedomain = _TimeDomain(t0, self.test_T)
        self.geomx = spacedomain
        self.geomt = timedomain
        return _GeometryXTime(spacedomain, timedomain)

    def generate_data(self, num_domain=100, num_boundary=20):
        geom = self.geometry()
        return geom.random_points(num_domain), geom.random_boundary_points(num_boundary)

    def generate_test_data(self, num_domain=100, num_boundary=20):
        geom = self.test_geometry()
        return geom.random_points(num_domain), geom.random_boundary_points(num_boundary)

    # ---- the same two samplers on the device (no counterpart in the reference: DeepXDE draws on the host with NumPy's global generator) ----
    def _points_device(self, n, boundary, seed, stream_id, radius, t0, T):
        torch = _lib.torch_cuda()
        d = self.n_input - 1
        out = torch.empty((int(n), d + 1), dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().scasml_geometry_points(int(seed), int(stream_id), int(n), d, -float(radius), float(radius), float(t0), float(T),
                                                      int(bool(boundary)), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def generate_data_device(self, num_domain=100, num_boundary=20, seed=1234):
        '''generate_data with the points drawn ON the device (Philox; float16-valued float64 CUDA tensors [n, d + 1]).  GP.GPsolver,
        GP.predict and the solvers' u_solve take them as they are; `.cpu().numpy().astype(np.float16)` gives the reference's host arrays.'''
        self.geometry()
        return (self._points_device(num_domain, False, seed, 0, self.radius, self.t0, self.T),
                self._points_device(num_boundary, True, seed, 1, self.radius, self.t0, self.T))

    def generate_test_data_device(self, num_domain=100, num_boundary=20, seed=42):
        self.test_geometry()
        return (self._points_device(num_domain, False, seed, 0, self.test_radius, self.t0, self.test_T),
                self._points_device(num_boundary, True, seed, 1, self.test_radius, self.t0, self.test_T))
