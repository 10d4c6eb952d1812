"""Statistical parity against the numbers the REAL reference committed (results/**/SimpleUniform.log,
RepeatedExperiment.log, InferenceScaling plots): the reference's own driver call patterns, run head-less through the
reference-shaped classes on the GPU (tools/reproduce_reference_tables.py).

The reference cannot run here (JAX / DeepXDE absent) and its collocation / test points cannot be re-drawn, so the tolerances are
statistical: the reference's 10-repetition standard deviation of the ScaSML error at d = 20 is 2.4e-3 on a mean of 6.9e-2, and the
GP error varies by ~5 % between training sets; the integer evaluation counters must match exactly."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def test_simple_uniform_repeated_and_inference_scaling_at_d20():
    import reproduce_reference_tables as R
    out = R.run_dim(20, reps=4)
    ref = R.REF
    for key, tol in (("gp", 0.12), ("mlp", 0.15), ("scasml", 0.12), ("scasml_rep", 0.10), ("scasml_rep_l1", 0.10),
                     ("mlp_fh", 0.15), ("scasml_fh", 0.12), ("scasml_fh_rep", 0.10), ("pde_std", 0.15)):
        want, got = ref[key][20], out[key]
        assert abs(got - want) <= tol * want, (key, got, want)
    # the surrogate's PDE residual is centred slightly below zero, like the reference's (-2.7e-3)
    assert -6e-3 < out["pde_mean"] < 0.0
    # ScaSML halves the surrogate's error (the paper's claim; reference: 0.1466 -> 0.0701)
    assert out["scasml"] < 0.6 * out["gp"] and out["scasml_fh"] < 0.6 * out["gp"]
    # InferenceScaling: improvement over min(GP, MLP) at rho = 1, 2, 3 and the x-coordinates of the reference's plot
    for got, want in zip(out["improvement"], ref["improvement"][20]):
        assert abs(got - want) < 8.0, (out["improvement"], ref["improvement"][20])
    assert tuple(out["counter"]) == ref["counter"]
    # evaluation_counter of the quadrature ScaSML at rho = 2 (83 per u_solve(2, 2))
    assert out["counter_scasml_rho2"] == 83
    assert np.isfinite(out["fit_s"]) and out["newton_steps"] <= 20
    # ConvergenceRate: slopes of the reference's committed plot (GP 0.37, SCaSML 0.57); ten refits of one GP object
    assert abs(out["slope_gp"] - ref["slope_gp"][20]) < 0.12, out["slope_gp"]
    assert abs(out["slope_scasml"] - ref["slope_scasml"][20]) < 0.16, out["slope_scasml"]
