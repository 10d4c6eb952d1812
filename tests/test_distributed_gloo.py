"""world_size-2 gloo test (CPU) of the multi-GPU decomposition: every rank owns the top-level sample units
u = rank + world*s, accumulates un-clipped weighted partial sums, one all_reduce(SUM) combines them, clip and the
float16 cast happen after the collective.  The compute engine here is the CPU oracle; the CUDA path
uses the same unit ownership rule (tests/test_gpu_parity.py::test_sample_sharding_partials_sum_to_full)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.equation import EquationOracle
from oracle.gp import GPOracle
from oracle.solvers import ScaSMLFullHistoryOracle, ScaSMLOracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(d=6, nd=30, nb=10):
    eq = EquationOracle(d + 1)
    gp = GPOracle(eq)
    dom, bdy = eq.generate_data(nd, nb, seed=1234)
    gp.GPsolver(dom, bdy, GN_steps=8)
    X = np.concatenate(eq.generate_test_data(7, 2, seed=42), axis=0)
    return eq, gp, X


def _worker(rank, world, port, variant, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eq, gp, X = _setup()
        if variant == "quadrature":
            s = ScaSMLOracle(eq, gp, cast=False)
            part = s.uz_solve(2, 2, X, shard=(rank, world))
        else:
            s = ScaSMLFullHistoryOracle(eq, gp, cast=False)
            part = s.uz_solve(2, None, X, 3, shard=(rank, world))
        t = torch.from_numpy(np.ascontiguousarray(part))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                 # the single collective of the path
        out = np.clip(t.numpy(), -eq.uncertainty, eq.uncertainty)
        # host-side bookkeeping must be rank-independent
        q.put((rank, out, s.evaluation_counter, s.key_counter))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["quadrature", "full_history"])
def test_two_rank_allreduce_reproduces_single_rank(variant):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eq, gp, X = _setup()
    if variant == "quadrature":
        ref = ScaSMLOracle(eq, gp, cast=False)
        ref.uz_solve(2, 2, X)
    else:
        ref = ScaSMLFullHistoryOracle(eq, gp, cast=False)
        ref.uz_solve(2, None, X, 3)
    for rank, out, counter, keys in results:
        np.testing.assert_allclose(out, ref.last_raw, rtol=1e-12, atol=1e-15)
        assert counter == ref.evaluation_counter and keys == ref.key_counter


def test_product_solver_detects_process_group():
    """PicardSolverBase._dist(): sharding is opt-in and only active inside an initialised process group."""
    from scasml_gp_b200.solvers._picard import PicardSolverBase
    s = PicardSolverBase()
    assert s._dist()[:2] == (0, 1)
    s.distributed = True
    assert s._dist()[:2] == (0, 1)                                # no process group -> single rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        assert s._dist()[:2] == (0, 1)
    finally:
        dist.destroy_process_group()
