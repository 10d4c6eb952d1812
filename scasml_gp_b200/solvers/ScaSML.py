"""ScaSML (quadrature) -- host mirror of the reference's ``solvers/ScaSML.py``."""
import numpy as np

from ._picard import PicardSolverBase


class ScaSML(PicardSolverBase):
    '''Multilevel Picard Iteration calibrated GP for high dimensional semilinear PDE'''
    variant = 0
    scasml = True

    def __init__(self, equation, GP):
        self._init_common(equation)          # solvers/ScaSML.py:8-27
        self.GP = GP

    def f(self, x_t, u_breve, z_breve):
        '''Defect generator (solvers/ScaSML.py:29-47).'''
        self.evaluation_counter += 1
        eq = self.equation
        u_hat = self.GP.predict(x_t)
        grad_u_hat_x = self.GP.compute_gradient(x_t, u_hat)[:, :-1]
        val1 = eq.f(x_t, np.asarray(u_breve) + u_hat, eq.sigma(x_t) * grad_u_hat_x + np.asarray(z_breve))
        val2 = eq.f(x_t, u_hat, eq.sigma(x_t) * grad_u_hat_x)
        return val1 - val2

    def g(self, x_t):
        '''Defect terminal condition (solvers/ScaSML.py:49-63).'''
        self.evaluation_counter += 1
        eq = self.equation
        u_hat = self.GP.predict(x_t)
        return (eq.g(x_t) - u_hat)[:, 0]

    def uz_solve(self, n, rho, x_t):
        '''(u_breve, z_breve), shape (batch, 1+d), float16 after clip (solvers/ScaSML.py:149-284).'''
        return self._uz(n, rho, x_t).astype(np.float16)

    def u_solve(self, n, rho, x_t):
        '''u_hat + u_breve, shape (batch, 1) (solvers/ScaSML.py:286-305).'''
        return self._u_solve(n, rho, x_t)
