"""MLP, full-history variant -- host mirror of ``solvers/MLP_full_history.py``."""
import numpy as np

from ._picard import PicardSolverBase
from .MLP import MLP


class MLP_full_history(PicardSolverBase):
    '''Multilevel Picard Iteration for high dimensional semilinear PDE'''
    variant = 1
    scasml = False

    def __init__(self, equation):
        self._init_common(equation)          # solvers/MLP_full_history.py:8-27

    f = MLP.f
    g = MLP.g

    def uz_solve(self, n, rho, x_t, M):
        return self._uz(n, rho, x_t, M).astype(np.float16)   # solvers/MLP_full_history.py:64-179

    def u_solve(self, n, rho, x_t, M=3):
        return self._u_solve(n, rho, x_t, M)                 # :181-196
