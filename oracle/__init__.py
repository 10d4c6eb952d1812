"""CPU oracle for the ScaSML correction hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy (float64) restatement of the reference algorithm
(Francis-Fan-create/SCaSML_GP: ``equations/equations.py``, ``models/GP.py``,
``solvers/{MLP,ScaSML}{,_full_history}.py``).  It exists to *check* the CUDA
product path and to provide the reported CPU baseline.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``scasml_gp_b200/`` does.

PARITY UNPINNED: the reference cannot be imported here (it needs jax, jaxlib and
deepxde, none installed, no network) and ships no unit tests, golden vectors or
known-answer fixtures for this path.  The oracle is therefore pinned against
  (1) a torch-autograd restatement of the reference's kernel functionals
      (``models/GP.py:28-180``) -- see ``tests/test_oracle_closed_forms.py``,
  (2) the integer tables / call counts / evaluation counters the reference's
      committed profiles and plots confirm, and
  (3) the statistical error aggregates of the reference's committed logs,
not against bit-level outputs of the reference.

Precision policy ("P-clean"): float64 internals; float16
rounding only on drawn normals/uniforms, on Gram entries (``models/GP.py:258``
semantics) and on public return values.  ``cast=True`` additionally rounds the
intermediate named outputs (f, g, predict, compute_gradient, compute_PDE_loss,
inner uz_solve returns) like the reference does; ``cast=False`` ("nocast") is
what the CUDA path implements internally.
"""
