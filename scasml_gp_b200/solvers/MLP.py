"""MLP (quadrature) -- host mirror of the reference's ``solvers/MLP.py``."""
import numpy as np

from ._picard import PicardSolverBase


class MLP(PicardSolverBase):
    '''Multilevel Picard Iteration for high dimensional semilinear PDE'''
    variant = 0
    scasml = False
    stale_delta = True       # solvers/MLP.py:201,249,270: the first z-update of a step re-uses the previous delta_t

    def __init__(self, equation):
        self._init_common(equation)          # solvers/MLP.py:8-25

    def f(self, x_t, u, z):
        return self.equation.f(x_t, u, z)    # solvers/MLP.py:27-40

    def g(self, x_t):
        return self.equation.g(x_t)[:, 0]    # solvers/MLP.py:42-55

    def uz_solve(self, n, rho, x_t):
        return self._uz(n, rho, x_t).astype(np.float16)      # solvers/MLP.py:141-274

    def u_solve(self, n, rho, x_t):
        return self._u_solve(n, rho, x_t)                    # solvers/MLP.py:276-288
