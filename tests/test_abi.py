"""The C-ABI library loads and exports every symbol include/scasml_b200.h declares (no compute without a GPU),
and the host-only plan entry point reproduces the reference's integer bookkeeping."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from scasml_gp_b200 import _lib
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "scasml_b200.h")).read()
    names = re.findall(r"SCASML_API\s+[\w\s\*]+?\b(scasml_\w+)\s*\(", hdr)
    assert len(names) >= 20
    cdll = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), n
    assert set(names) == set(lib._SIGNATURES), set(names) ^ set(lib._SIGNATURES)
    assert lib.load().scasml_abi_version() == 1
    # the product library exports nothing else: test hooks and micro-benchmarks live in the debug build only
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert exported == set(names), exported ^ set(names)


def test_debug_header_symbols_are_exported_by_the_debug_build_only(lib):
    hdr = open(os.path.join(ROOT, "include", "scasml_b200_debug.h")).read()
    names = re.findall(r"SCASML_API\s+[\w\s\*]+?\b(scasml_\w+)\s*\(", hdr)
    assert set(names) == set(lib._DEBUG_SIGNATURES), set(names) ^ set(lib._DEBUG_SIGNATURES)
    dbg = C.CDLL(lib.LIB_DBG_PATH)
    prod = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dbg, n), n
        assert not hasattr(prod, n), n


def test_struct_layout_matches_header(lib):
    # sizes are static_assert'ed against the C++ structs in csrc/abi.cu; here: the Python mirror vs the header
    assert C.sizeof(lib.PicardStats) == 12 * 8
    ints = 6 + 8 + 8 + 9
    assert C.sizeof(lib.PicardParams) == ((ints * 4 + 7) // 8) * 8 + 2 * 64 * 8 + 4 * 8 + 6 * 4 + 8 + 8


class _Eq:
    n_output, uncertainty, norm_estimation, T, t0 = 1, 0.1, 1, 0.5, 0

    def __init__(self, d):
        self.n_input = d + 1

    def sigma(self, x=0):
        return 0.25

    def mu(self, x=0):
        return -1 / (self.n_input - 1) - 0.25 ** 2 / 2

    def geometry(self):
        pass


def test_plan_reproduces_reference_bookkeeping(lib):
    from scasml_gp_b200.solvers.MLP import MLP
    from scasml_gp_b200.solvers.MLP_full_history import MLP_full_history
    from scasml_gp_b200.solvers.ScaSML import ScaSML
    from scasml_gp_b200.solvers.ScaSML_full_history import ScaSML_full_history
    want = {1: (10, 5, 10, 7), 2: (83, 46, 54, 33), 3: (549, 372, 246, 127), 4: (3659, 2714, 1034, 449)}
    sp_q = {2: 46, 3: 1090, 4: 13426}
    sp_fh = {1: 9, 2: 60, 3: 390, 4: 2523}
    calls = {2: 1 + 3, 3: 1 + 3 + 15, 4: 1 + 3 + 16 + 77}          # Python-level uz_solve calls of level >= 1
    for n, (c_s, c_m, c_sf, c_mf) in want.items():
        _, st = ScaSML(_Eq(4), None).plan(n, n, 1)
        assert st.eval_counter == c_s
        if n in sp_q:
            assert st.sample_points == sp_q[n] and st.n_calls == calls[n]
        assert MLP(_Eq(4)).plan(n, n, 1)[1].eval_counter == c_m
        _, sf = ScaSML_full_history(_Eq(4), None).plan(n, None, 1, M=3)
        assert sf.eval_counter == c_sf and sf.sample_points == sp_fh[n]
        assert MLP_full_history(_Eq(4)).plan(n, None, 1, M=3)[1].eval_counter == c_mf
    # executed work excludes the discarded level-0 terminals: 13 426 - 4 844 at n = rho = 4
    assert ScaSML(_Eq(4), None).plan(4, 4, 1)[1].executed_points == 13426 - 4844


def test_plan_sharding_partitions_the_units(lib):
    from scasml_gp_b200.solvers.ScaSML import ScaSML
    B = 13
    _, full = ScaSML(_Eq(7), None).plan(3, 3, B)
    for world in (2, 3, 8):
        tot = sum(ScaSML(_Eq(7), None).plan(3, 3, B, rank=r, world=world)[1].executed_points for r in range(world))
        assert tot == full.executed_points
        keys = {ScaSML(_Eq(7), None).plan(3, 3, B, rank=r, world=world)[1].keys_used for r in range(world)}
        assert keys == {full.keys_used}


def test_tables_match_oracle(lib):
    from oracle.tables import approx_parameters as ap_o
    from scasml_gp_b200.solvers._picard import approx_parameters as ap_p
    for rho in (1, 2, 3, 4):
        for quad, gl in (("reference", False), ("gauss_legendre", True)):
            a, b = ap_p(rho, 0.5, quad), ap_o(rho, 0.5, true_gl=gl)
            for x, y in zip(a, b):
                assert np.array_equal(x, y, equal_nan=True)
    assert np.array_equal(lib.normal_half_table().view(np.uint16),
                          __import__("oracle.rng", fromlist=["x"]).normal_half_table().view(np.uint16))


def test_missing_library_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libscasml_b200.so")
    with pytest.raises(lib.ScasmlError):
        lib.load()


def test_no_cuda_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
    with pytest.raises(lib.ScasmlError):
        Grad_Dependent_Nonlinear(5).g(np.zeros((2, 5)))
