"""ScaSML, full-history variant -- host mirror of ``solvers/ScaSML_full_history.py``."""
import numpy as np

from ._picard import PicardSolverBase
from .ScaSML import ScaSML


class ScaSML_full_history(PicardSolverBase):
    '''Multilevel Picard Iteration calibrated GP for high dimensional semilinear PDE'''
    variant = 1
    scasml = True

    def __init__(self, equation, GP):
        self._init_common(equation)          # solvers/ScaSML_full_history.py:8-27
        self.GP = GP

    f = ScaSML.f                             # solvers/ScaSML_full_history.py:29-50 (same defect generator)
    g = ScaSML.g                             # :52-72

    def uz_solve(self, n, rho, x_t, M):
        # the reference returns the clipped array without a cast (:199); its dtype flow is float16
        return self._uz(n, rho, x_t, M).astype(np.float16)

    def u_solve(self, n, rho, x_t, M=3):
        return self._u_solve(n, rho, x_t, M)                 # :201-221
