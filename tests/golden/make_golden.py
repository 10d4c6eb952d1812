"""Regenerate tests/golden/oracle_vectors_d6.npz: small input/output vectors of the hot path produced by the CPU oracle.

The reference itself cannot be imported in this environment (jax / jaxlib / deepxde are not installable), so these vectors
come from `oracle/` (float64 NumPy restatement of the reference, "parity unpinned", see oracle/__init__.py and DESIGN.md);
what pins the oracle against the reference is tests/golden/reference_known_answers.json and the autograd restatement in
tests/ref_autograd.py.  The vectors serve two purposes: (1) `-m "not gpu"` tests fail if the oracle drifts, (2) `-m gpu`
tests compare the CUDA path with committed numbers on a box where nothing else is available.

    python tests/golden/make_golden.py        (run from the repo root; deterministic)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.equation import EquationOracle                                    # noqa: E402
from oracle.gp import GPOracle                                                # noqa: E402
from oracle.solvers import MLPOracle, ScaSMLFullHistoryOracle, ScaSMLOracle   # noqa: E402


def build():
    d, nd, nb = 6, 40, 12
    eq = EquationOracle(d + 1)
    dom, bdy = eq.generate_data(nd, nb, seed=1234)
    idx = np.random.default_rng(7).choice(d, 5, replace=False)
    sol0 = np.random.default_rng(0).standard_normal(3 * nd) * 1e-3
    gp = GPOracle(eq, idx_set=idx)
    gp.GPsolver(dom, bdy, sol0=sol0)
    X = np.concatenate(eq.generate_test_data(13, 4, seed=42), axis=0).astype(np.float64)
    X[:3] += 0.0123456789                       # not float16-representable
    eps, u, dv, lp, dt = gp.pde_terms_raw(X)
    out = dict(d=d, idx=idx, dom=dom, bdy=bdy, sol0=sol0, alpha=gp.right_vector[:, 0], loss_history=np.asarray(gp.loss_history),
               X=X, u=u, div=dv, lap=lp, dt=dt, eps=eps, grad=gp.gradient_raw(X), g=eq.g(X, cast=False)[:, 0],
               exact=eq.exact_solution(X).astype(np.float64))
    s = ScaSMLOracle(eq, gp, cast=False)
    s.uz_solve(2, 2, X)
    out["scasml_uz_n2"] = s.last_raw
    out["scasml_counter_n2"] = s.evaluation_counter
    s = ScaSMLOracle(eq, gp, cast=False, true_gl=True)
    s.uz_solve(3, 3, X)
    out["scasml_gl_uz_n3"] = s.last_raw
    s = ScaSMLFullHistoryOracle(eq, gp, cast=False)
    s.uz_solve(2, None, X, M=3)
    out["scasml_fh_uz_n2_M3"] = s.last_raw
    s = MLPOracle(eq, cast=False)
    s.uz_solve(2, 2, X)
    out["mlp_uz_n2"] = s.last_raw
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors_d6.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
