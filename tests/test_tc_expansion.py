"""CPU check of the separable (two-GEMM) expansion the tcgen05 route evaluates: the column table and per-centre
coefficients of tests/tc_expansion_ref.py (mirrored by coef_of() in csrc/gp_eval_tc.cu) against the oracle's closed forms
(models/GP.py:630-687, 746-769), plus the predicted accuracy of the f16 hi/lo operand splits."""
import numpy as np
import pytest

from oracle.equation import EquationOracle
from oracle.gp import GPOracle
from tests import tc_expansion_ref as E


def _gp(d, nd, nb, seed=0):
    eq = EquationOracle(d + 1)
    dom, bdy = eq.generate_data(nd, nb, seed=1234)
    gp = GPOracle(eq, idx_set=np.random.default_rng(seed).choice(d, 5, replace=False))
    gp.gram(dom, bdy)
    gp.right_vector = (np.random.default_rng(seed + 1).standard_normal(gp.phi_dim) * 0.3)[:, None]
    X = np.concatenate(eq.generate_test_data(40, 10, seed=7), axis=0).astype(np.float64)
    X[:9] += 0.0123456789
    return gp, X


@pytest.mark.parametrize("d", [6, 20])
def test_expansion_matches_closed_forms(d):
    gp, X = _gp(d, 30, 9)
    want = {E.OUT_U: gp._row_dot(X, "id"), E.OUT_G: gp._row_dot(X, "div"), E.OUT_L: gp._row_dot(X, "lap"), E.OUT_T: gp._row_dot(X, "dt")}
    for mode, outs in ((E.MODE_U, [E.OUT_U]), (E.MODE_UG, [E.OUT_U, E.OUT_G]), (E.MODE_PDE, [E.OUT_U, E.OUT_G, E.OUT_L, E.OUT_T])):
        got = E.evaluate(mode, gp, X)
        for o in outs:
            scale = max(1.0, np.abs(want[o]).max())
            assert np.max(np.abs(got[o] - want[o])) < 1e-10 * scale, (mode, o)


def test_column_counts_fit_the_tensor_memory_budget():
    # padded to multiples of 16 (UMMA N granularity at M = 128): PDE 48 + 48 + 32 = 128 accumulator columns
    n = {(m, c): len(E.columns(m, c)) for m in (0, 1, 2) for c in (0, 1, 2)}
    assert n[(0, 0)] == 3 and n[(0, 1)] == 7 and n[(0, 2)] == 0
    assert n[(1, 0)] == 8 and n[(1, 1)] == 21
    assert n[(2, 0)] == 41 and n[(2, 1)] == 35 and n[(2, 2)] == 21


def test_split_operand_emulation_is_accurate():
    gp, X = _gp(20, 60, 12)
    exact = E.evaluate(E.MODE_PDE, gp, X)
    emu = E.evaluate(E.MODE_PDE, gp, X, emulate=True)
    for o in exact:
        scale = max(1.0, np.abs(exact[o]).max())
        assert np.max(np.abs(emu[o] - exact[o])) < 2e-5 * scale, o
        assert np.sqrt(np.mean((emu[o] - exact[o]) ** 2)) < 5e-7 * np.sqrt(np.mean(exact[o] ** 2)), o
