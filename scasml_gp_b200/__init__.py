"""scasml_gp_b200 -- B200-native (sm_100a) implementation of SCaSML_GP's ScaSML correction hot path.

The package mirrors the reference's Python API for that path only:
    equations.equations.Grad_Dependent_Nonlinear      (reference equations/equations.py:232-417)
    models.GP.GP_Grad_Dependent_Nonlinear             (reference models/GP.py)
    solvers.{MLP, ScaSML, MLP_full_history, ScaSML_full_history}
and routes every computation through the C ABI of ``libscasml_b200.so`` (include/scasml_b200.h).
There is no CPU fallback: without the compiled library and a CUDA device the classes raise.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
