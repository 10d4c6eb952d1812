#!/usr/bin/env python
"""bench.py -- ScaSML correction throughput (sample-points/s) on 1..8 B200, next to the CPU oracle.

A "step" is one pass of the hot path over one batch of synthetic test points: the flattened multilevel-Picard
correction (Philox sampling, fused surrogate evaluation, reductions) plus the final u_hat + u_breve.  The GP fit is
done once before the timed region and reported as ``fit`` (time, FP64 roofline) and inside ``e2e_with_fit``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]
  torchrun ... bench.py --gpus N ...     (one rank per GPU; samples sharded, one all-reduce per step)

JSON keys follow the driver's contract (value, e2e, roofline, cpu_baseline, clocks, gpu_launches ...) plus
  accuracy     rel-L2 / L1 error vs the exact solution (tests/SimpleUniform.py:110-136) of the product and of the CPU oracle
               on the same test points and increments, and their relative difference (north-star criterion: <= 1e-6)
  nrank_check  (N > 1) the all-reduced N-rank result against a single-GPU solve of the same test points
  strong       (N > 1) the same total batch as the 1-GPU run, sharded over N ranks (strong scaling), with the 1-GPU time of
               that batch measured in the same run
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # cpu_points: test points of the bounded CPU sample (cpu_baseline leg and every step of --impl reference)
    "C2": dict(d=20, nd=1000, nb=200, n=3, rho=3, variant="quadrature", M=None, cpu_points=48,
               desc="Grad_Dependent_Nonlinear d=20, ScaSML n=rho=3, GP 1000+200 collocation points"),
    "C3": dict(d=100, nd=1000, nb=200, n=4, rho=4, variant="quadrature", M=None, cpu_points=6,
               desc="Grad_Dependent_Nonlinear d=100, ScaSML n=rho=4, GP 1000+200 collocation points (phi=4200)"),
    "C5": dict(d=1000, nd=4000, nb=800, n=4, rho=4, variant="quadrature", M=None, cpu_points=1,
               desc="Grad_Dependent_Nonlinear d=1000, ScaSML n=rho=4, GP 4000+800 collocation points (phi=16800); K-streamed tcgen05 kernel"),
    "C4": dict(d=60, nd=1000, nb=200, n=4, rho=None, variant="full_history", M=3, cpu_points=24,
               desc="Grad_Dependent_Nonlinear d=60, ScaSML_full_history n=4 M=3"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=1200, help="test points per GPU (reference drivers use 1 200)")
    ap.add_argument("--route", default=os.environ.get("SCASML_ROUTE", "auto"), choices=["auto", "f64", "tc"])
    ap.add_argument("--quadrature", default="gauss_legendre", choices=["reference", "gauss_legendre"],
                    help="n=4 with the reference's own lgwt tables is NaN everywhere (SURVEY quirk 1)")
    ap.add_argument("--cpu-points", type=int, default=0, help="test points of the bounded CPU sample (0: per-config default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    return ap.parse_args()


def gen_points(d, nd, nb, B_total):
    """Synthetic collocation / test points (same law as the reference's DeepXDE sampling), product-side only."""
    rng = np.random.default_rng(1234)

    def pts(n, boundary, r):
        x = r.random((n, d))
        if boundary:
            dim = r.integers(0, d, size=n)
            x[np.arange(n), dim] = np.round(x[np.arange(n), dim])
        x = x - 0.5
        t = r.permutation(r.random((n, 1)) * 0.5)
        return np.hstack([x, t]).astype(np.float16).astype(np.float64)
    dom, bdy = pts(nd, False, rng), pts(nb, True, rng)
    r2 = np.random.default_rng(42)
    nbt = B_total // 6
    X = np.concatenate([pts(B_total - nbt, False, r2), pts(nbt, True, r2)], axis=0)
    return dom, bdy, X


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        # NVML in-process (nvidia_ml_py): ~10 ms per sample, so even a half-second timed region gets dozens of samples;
        # nvidia-smi as a subprocess (~0.3 s per call) is the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            while not self.stop_flag:
                r = int(get_reasons(h))
                self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx)]
                                 + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.01)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.03)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_run(cfg, dom, bdy, Xs, alpha, idx, quadrature):
    """One pass of the CPU oracle (NumPy + OpenBLAS, all host threads) over the test points Xs.  Returns (seconds, solver)."""
    from oracle.equation import EquationOracle
    from oracle.gp import GPOracle
    from oracle.solvers import ScaSMLFullHistoryOracle, ScaSMLOracle
    eq_o = EquationOracle(cfg["d"] + 1)
    gp_o = GPOracle(eq_o, idx_set=idx)
    gp_o.x_t_domain, gp_o.x_t_boundary = dom, bdy
    gp_o.N_domain, gp_o.N_boundary = len(dom), len(bdy)
    gp_o.right_vector = alpha.reshape(-1, 1)
    fh = cfg["variant"] == "full_history"
    s = (ScaSMLFullHistoryOracle if fh else ScaSMLOracle)(eq_o, gp_o, cast=False, true_gl=(quadrature != "reference"))
    t0 = time.perf_counter()
    if fh:
        s.u_solve(cfg["n"], None, Xs, M=cfg["M"])
    else:
        s.u_solve(cfg["n"], cfg["rho"], Xs)
    return time.perf_counter() - t0, s


def exact_solution(X):
    """u = 1 - 1 / (1 + exp(t + sum x))  (equations/equations.py:306-323), float64."""
    return 1.0 - 1.0 / (1.0 + np.exp(X.sum(axis=1)))


def errors_vs_exact(sol, exact):
    """rel-L2 and mean-L1 error with the NaN masking of tests/SimpleUniform.py:110-136."""
    sol, exact = np.asarray(sol, dtype=np.float64).ravel(), np.asarray(exact, dtype=np.float64).ravel()
    m = ~(np.isnan(sol) | np.isnan(exact))
    if not m.any():
        return float("nan"), float("nan")
    return float(np.linalg.norm(sol[m] - exact[m]) / np.linalg.norm(exact[m])), float(np.mean(np.abs(sol[m] - exact[m])))


def measured_traffic(route):
    """DRAM bytes per evaluated point of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch / its points).  Stale captures (the kernel
    source changed since) give None instead of a number that no longer describes the code."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        src = open(os.path.join(ROOT, rec["kernel_source"]), "rb").read()
        if route != "tc" or hashlib.sha256(src).hexdigest()[:16] != rec["kernel_source_sha16"]:
            return None, "profiles/ncu_traffic.json is stale for this build (kernel source changed since the capture)" if route == "tc" else "no capture for this route"
        return float(rec["dram_bytes_per_point"]), f"{rec['dram_bytes_per_point']} B/point from {rec['capture']} x points per launch"
    except Exception as e:                                   # missing file: say so
        return None, f"no committed capture ({type(e).__name__})"


def fit_flops(d, nd, nb, steps, lu):
    """Algorithmic FP64 flops of GPsolver: Gram distance contractions, Cholesky + explicit inverse of K + nugget I
    (phi^3/3 + 2 phi^3/3), one factorisation of the (3 N_d)^2 Newton system per step (Cholesky n^3/3, or LU 2 n^3/3)."""
    phi, n3 = 4 * nd + nb, 3 * nd
    return 3 * 2 * (d + 1) * (nd + nb) ** 2 + phi ** 3 + steps * (2 if lu else 1) * n3 ** 3 / 3


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d, n, rho, M = cfg["d"], cfg["n"], cfg["rho"], cfg["M"]
    fh = cfg["variant"] == "full_history"
    B_total = args.batch * max(world, 1)
    dom, bdy, X = gen_points(d, cfg["nd"], cfg["nb"], B_total)
    idx = np.random.default_rng(0).choice(d, 5, replace=False)
    cpu_points = args.cpu_points or cfg["cpu_points"]
    # identical in both arms (the driver compares it byte for byte)
    config = {"workload": cfg["desc"], "config": args.config, "test_points_per_gpu": args.batch, "test_points_total": B_total,
              "parallelism": f"sample-sharded x{world}" if world > 1 else "single GPU",
              "quadrature_tables": args.quadrature, "l2": "level point buffers (GBs) exceed L2 every step",
              "sample_point_unit": "executed sample points (level-0 terminals the reference draws and discards are elided)"}

    # ------------------------------------------------------------------ reference arm: CPU oracle ----------
    if args.impl == "reference":
        if rank != 0:
            return
        # fit on CPU (a bounded sample of the path is what is timed; the fit is setup)
        from oracle.equation import EquationOracle
        from oracle.gp import GPOracle
        eq_o = EquationOracle(d + 1)
        gp_o = GPOracle(eq_o, idx_set=idx)
        t0 = time.perf_counter()
        gp_o.GPsolver(dom, bdy)
        fit_s = time.perf_counter() - t0
        alpha = gp_o.right_vector[:, 0]
        rates, dts = [], []
        for it in range(args.warmup + args.steps):
            dt, s = cpu_oracle_run(cfg, dom, bdy, X[:cpu_points], alpha, idx, args.quadrature)
            if it >= args.warmup:
                rates.append(_exec_sp(s) / dt)
                dts.append(dt)
        val = float(np.mean(rates))
        line = {"impl": "reference", "metric": "ScaSML correction sample-points/s", "value": val, "unit": "sample-points/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "sample-points/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{cpu_points} test points of the same workload per step (NumPy+OpenBLAS oracle, all host threads)",
                                 "fit_s": fit_s},
                "e2e": {"value": val, "unit": "sample-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm ----------------------------
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from scasml_gp_b200 import _lib
    from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
    from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
    from scasml_gp_b200.solvers.ScaSML import ScaSML
    from scasml_gp_b200.solvers.ScaSML_full_history import ScaSML_full_history

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eq = Grad_Dependent_Nonlinear(d + 1)
    gp = GP_Grad_Dependent_Nonlinear(eq, idx_set=idx)
    route = args.route
    if route == "auto":     # the tcgen05 route is the product path wherever it applies (d <= 1022); FP64 is the parity anchor
        route = "tc" if d + 2 <= 1024 else "f64"
    sroute = _lib.ROUTE_TC if route == "tc" else _lib.ROUTE_F64   # route of the sampled points; the top-level u_hat(x) of u_solve
                                                                  # stays on GP.route = FP64, the public API's default
    gp.GPsolver(dom, bdy)                                    # first fit: allocations, module load
    torch.cuda.synchronize()
    fit_times = []
    for _ in range(1 if d > 300 else 3):                     # median of three fits (a single wall-clock sample now and then caught a 10x host hiccup)
        t0 = time.perf_counter()
        gp.GPsolver(dom, bdy)
        torch.cuda.synchronize()
        fit_times.append(1e3 * (time.perf_counter() - t0))
    fit_ms = sorted(fit_times)[len(fit_times) // 2]

    Solver = ScaSML_full_history if fh else ScaSML

    def new_solver(distributed):
        s = Solver(eq, gp)
        s.route = sroute
        s.quadrature = args.quadrature
        s.distributed = distributed
        s.use_graph = not args.no_graph
        return s

    solver = new_solver(world > 1)
    x_dev = _lib.to_device(X)
    torch.cuda.synchronize()

    def step_device(s=solver, xd=x_dev):
        # the product path of u_solve on device-resident inputs: correction + top-level u_hat(x); sharded runs evaluate u_hat on the rows
        # r = rank (mod world) only and it travels in the same all-reduce (every rank used to evaluate all N x 1 200 rows here)
        uz, uh = s._uz_device(n, rho, xd, M, with_u_hat=True)
        return uz[:, 0] + uh

    def step_e2e(s=solver, Xh=X):
        return s.u_solve(n, rho, Xh, M) if fh else s.u_solve(n, rho, Xh)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        step_device()
    # device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    sampler.stop_flag = True                  # clocks are sampled during the timed region above; NVML queries contend with kernel launches for the
                                              # driver, which the short steps of the small configurations feel in the end-to-end loop below
    stats = dict(solver.last_stats)
    # kernel-group timing (separate, event-bracketed pass so the timed region above stays sync-free)
    solver.timing = True
    step_device()
    tstats = dict(solver.last_stats)
    solver.timing = False
    # end-to-end: host buffers in, host result out, through the reference-shaped API
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop_flag = True
    full_l2, full_l1 = errors_vs_exact(solver.last_raw_u, exact_solution(X))

    # strong scaling: the 1-GPU batch (args.batch test points in total) sharded over all ranks, and the same batch on rank 0 alone
    strong = None
    nrank = None
    if dist is not None:
        if not args.no_strong:
            Xs = X[:args.batch]
            xs_dev = _lib.to_device(Xs)
            sh = new_solver(True)
            for _ in range(2):
                step_device(sh, xs_dev)
            ms_sh = timed(lambda: step_device(sh, xs_dev), args.steps)
            step_e2e(sh, Xs)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e(sh, Xs)
            barrier()
            e2e_sh = 1e3 * (time.perf_counter() - t0)
            one = new_solver(False)
            ms_one = float("nan")
            if rank == 0:                                     # the other ranks wait at the barrier inside timed()
                for _ in range(2):
                    step_device(one, xs_dev)
            ms_one = timed((lambda: step_device(one, xs_dev)) if rank == 0 else (lambda: None), args.steps)
            strong = [ms_sh, e2e_sh, ms_one]
        # N ranks = 1 rank, on the hardware: fresh solvers (same split counter), first test points
        Xc = X[:16]
        a = new_solver(True)
        b = new_solver(False)
        step_e2e(a, Xc)
        step_e2e(b, Xc)
        same_nan = bool(np.array_equal(np.isnan(a.last_raw), np.isnan(b.last_raw)))
        diff = float(np.nanmax(np.abs(a.last_raw - b.last_raw))) if np.isfinite(a.last_raw).any() else 0.0
        nrank = [diff, 0.0 if same_nan else 1.0]              # second entry: NaN patterns differ (max over ranks)

    if dist is not None:
        t = torch.tensor([ms, e2e_ms] + (strong or [0.0, 0.0, 0.0]) + nrank, dtype=torch.float64, device="cuda")
        t[4] = torch.nan_to_num(t[4], nan=0.0)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        if strong:
            strong = [float(t[2]), float(t[3]), float(t[4])]
        nrank = [float(t[5]), float(t[6])]
        ex = torch.tensor([float(stats["executed_points"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(ex, op=dist.ReduceOp.SUM)
        executed_total = float(ex[0])
    else:
        executed_total = float(stats["executed_points"])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if nrank is not None:
        assert nrank[0] < 1e-12 and nrank[1] == 0.0, f"N-rank result differs from the single-GPU result: {nrank}"

    ms_per_step = ms / args.steps
    value = executed_total / (ms_per_step * 1e-3)
    ref_equiv = stats["sample_points"] * B_total / (ms_per_step * 1e-3)
    e2e_value = executed_total / (e2e_ms / args.steps * 1e-3)
    finite = float(np.mean(np.isfinite(res.astype(np.float64))))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # FP64 pipe peak, measured here with a cuBLAS fp64 GEMM (MEASURED_PEAKS.json has no fp64 figure): denominator of the fit
    # and of the FP64 evaluation route
    a64 = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    torch.matmul(a64, a64)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        torch.matmul(a64, a64)
    g1.record()
    torch.cuda.synchronize()
    fp64_peak = 3 * 2 * 4096 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
    del a64

    eval_s = tstats["eval_time_ns"] * 1e-9
    achieved_tf = tstats["eval_flops"] / max(eval_s, 1e-12) / 1e12
    if route == "tc":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_note = "measured cuBLAS bf16 sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PFLOP/s sustained"
        flop_note = ("kind::f16 tcgen05, two chained GEMMs, operands split hi+lo in f16 (2 distance passes, 3 coefficient products); "
                     "credited flops are the algorithmic 2(d+1) per pair-distance only")
    else:
        peak = fp64_peak
        peak_note = "FP64 route: cuBLAS fp64 4096^3 GEMM measured in this run (MEASURED_PEAKS.json has no fp64 figure)"
        flop_note = "FP64 FMA pipe (SIMT contraction); credited flops are the algorithmic 2(d+1) per pair-distance"
    bpp, traffic_note = measured_traffic(route)
    traffic = bpp * tstats["eval_points_total"] / max(tstats["eval_launches"], 1) if bpp is not None else None
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak,
                "traffic": traffic, "traffic_note": "bytes per evaluation launch (mean): " + traffic_note,
                "kernel": "fused surrogate evaluation (gp_eval*.cu)", "peak_source": peak_note,
                "note": flop_note, "eval_share_of_step": tstats["eval_time_ns"] / max(
                    tstats["eval_time_ns"] + tstats["sample_time_ns"] + tstats["reduce_time_ns"], 1),
                "eval_ms": 1e-6 * tstats["eval_time_ns"], "sample_ms": 1e-6 * tstats["sample_time_ns"],
                "reduce_ms": 1e-6 * tstats["reduce_time_ns"], "eval_launches": tstats["eval_launches"]}

    # HBM-bound kernels of the step (samplers write, reduction reads the level point buffers: (d+1) * 8 B per sample point)
    hbm_peak = peaks.get("hbm_gbs", 6554.0) if peaks else 6554.0
    pt_bytes = (d + 1) * 8.0 * float(stats["executed_points"])
    # what the samplers write per point: the FP64 row + its global id, and on the resident-operand tcgen05 route (d + 2 <= 128) the operand record
    # (16 nstep words + 16 B) and the 96-byte feature block of the evaluation kernel (csrc/gp_tc.cuh)
    nstep = next((k for k in (2, 4, 7, 8) if k >= -(-(d + 2) // 16)), 0)
    samp_bpp = (d + 1) * 8 + 8 + ((nstep * 64 + 16 + 96) if (route == "tc" and nstep) else 0)
    samp_bytes = float(samp_bpp) * float(stats["executed_points"])
    hbm = {"unit": "GB/s", "peak": hbm_peak, "bytes_per_point": (d + 1) * 8, "sampler_bytes_per_point": samp_bpp,
           "sampler": {"achieved": samp_bytes / max(tstats["sample_time_ns"], 1), "frac": samp_bytes / max(tstats["sample_time_ns"], 1) / hbm_peak},
           "reduction": {"achieved": pt_bytes / max(tstats["reduce_time_ns"], 1), "frac": pt_bytes / max(tstats["reduce_time_ns"], 1) / hbm_peak}}
    roofline["hbm_kernels"] = hbm

    # the fit (models/GP.py:487-604): FP64 tensor-core (DMMA) products, blocked Cholesky / LU
    ff = fit_flops(d, cfg["nd"], cfg["nb"], gp.newton_steps, False)   # credited at the Cholesky count even where the pivoted LU runs
    fit = {"ms": fit_ms, "newton_steps": gp.newton_steps, "flops": ff, "achieved": ff / (fit_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
           "peak": fp64_peak, "frac": ff / (fit_ms * 1e-3) / 1e12 / fp64_peak, "bound": "fp64 pipe",
           "peak_source": "cuBLAS fp64 4096^3 GEMM measured in this run"}
    e2e_fit_value = executed_total / (e2e_ms / args.steps * 1e-3 + fit_ms * 1e-3)

    # accuracy: product vs CPU oracle on the same test points and increments (fresh solvers: same split counter), vs the exact solution
    cpu = None
    accuracy = {"rel_l2": full_l2, "l1": full_l1, "points": int(len(X)),
                "note": "error of u_hat + u_breve vs the exact solution over all test points of the last end-to-end step"}
    if not args.no_cpu_baseline and world == 1:
        Xc = X[:cpu_points]
        dt, so = cpu_oracle_run(cfg, dom, bdy, Xc, gp.right_vector[:, 0], idx, args.quadrature)
        cpu = {"value": _exec_sp(so) / dt, "unit": "sample-points/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{cpu_points} test points of the same workload, {dt:.1f} s (NumPy+OpenBLAS oracle, all host threads)"}
        sp = new_solver(False)
        step_e2e(sp, Xc)
        ex = exact_solution(Xc)
        l2_p, l1_p = errors_vs_exact(sp.last_raw_u, ex)
        l2_o, l1_o = errors_vs_exact(so.last_raw_u, ex)
        accuracy.update({"oracle_points": int(cpu_points), "rel_l2_product": l2_p, "l1_product": l1_p, "rel_l2_oracle": l2_o,
                         "l1_oracle": l1_o, "rel_diff": abs(l2_p - l2_o) / l2_o if l2_o > 0 else float("nan"),
                         "rel_diff_l1": abs(l1_p - l1_o) / l1_o if l1_o > 0 else float("nan"),
                         "max_abs_diff_u": float(np.nanmax(np.abs(sp.last_raw_u - so.last_raw_u))),
                         "counters_equal": bool(sp.evaluation_counter == so.evaluation_counter)})

    D = d + 1
    line = {"metric": "ScaSML correction sample-points/s", "value": value, "unit": "sample-points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if route == "f64" else "f16 (split operands, fp32 tensor-core accumulate, f64 final contraction)",
            "data": "synthetic", "config": config, "route": route, "cuda_graph": bool(getattr(solver, "graph_replays", 0) > 0),
            "reference_equivalent_value": ref_equiv, "fit_ms": fit_ms, "newton_steps": gp.newton_steps,
            "finite_fraction": finite,
            "e2e": {"value": e2e_value, "unit": "sample-points/s", "h2d_bytes_per_step": B_total * D * 8,
                    "d2h_bytes_per_step": 2 * B_total * 8, "ms_per_step": e2e_ms / args.steps},
            "e2e_with_fit": {"value": e2e_fit_value, "unit": "sample-points/s",
                             "note": "one GP fit + one end-to-end correction pass over the batch"},
            "gpu_launches": int((stats["launches"] + 1) * args.steps),
            "roofline": roofline, "fit": fit, "accuracy": accuracy, "cpu_baseline": cpu, "clocks": sampler.summary()}
    if strong is not None:
        # executed sample points of the strong batch = per-test-point executed points x args.batch (independent of the sharding)
        per_tp = executed_total / B_total
        line["strong"] = {"test_points_total": args.batch, "ms_per_step": strong[0] / args.steps,
                          "value": per_tp * args.batch / (strong[0] / args.steps * 1e-3), "unit": "sample-points/s",
                          "e2e_ms_per_step": strong[1] / args.steps, "n1_ms_per_step": strong[2] / args.steps,
                          "speedup_vs_n1": strong[2] / strong[0], "efficiency_vs_n1": strong[2] / strong[0] / world,
                          "note": "same total batch as the 1-GPU run, sample units sharded over the ranks; n1 = that batch on rank 0 alone, same run"}
    if nrank is not None:
        line["nrank_check"] = {"points": 16, "max_abs_diff": nrank[0], "nan_pattern_equal": bool(nrank[1] == 0.0),
                               "note": "all-reduced N-rank (u, z) vs a single-GPU solve of the same test points"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def _exec_sp(solver):
    """Executed sample points of an oracle run: everything drawn minus the discarded level-0 terminals."""
    return float(getattr(solver, "sample_points_executed", solver.sample_points))


if __name__ == "__main__":
    main()
