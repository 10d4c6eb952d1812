#!/usr/bin/env python
"""bench.py -- ScaSML correction throughput (sample-points/s) on 1..8 B200, next to the CPU oracle.

A "step" is one pass of the hot path over one batch of synthetic test points: the flattened multilevel-Picard
correction (Philox sampling, fused surrogate evaluation, reductions) plus the final u_hat + u_breve.  The GP fit is
done once before the timed region and reported as ``fit_ms``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...     (one rank per GPU; samples sharded, one all-reduce per step)

JSON keys follow the driver's contract (value, e2e, roofline, cpu_baseline, clocks, gpu_launches ...).
"""
import argparse
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
    # name: (d, N_d, N_b, n, rho, variant, M)
    "C2": dict(d=20, nd=1000, nb=200, n=3, rho=3, variant="quadrature", M=None,
               desc="Grad_Dependent_Nonlinear d=20, ScaSML n=rho=3, GP 1000+200 collocation points"),
    "C3": dict(d=100, nd=1000, nb=200, n=4, rho=4, variant="quadrature", M=None,
               desc="Grad_Dependent_Nonlinear d=100, ScaSML n=rho=4, GP 1000+200 collocation points (phi=4200)"),
    "C5": dict(d=1000, nd=4000, nb=800, n=4, rho=4, variant="quadrature", M=None,
               desc="Grad_Dependent_Nonlinear d=1000, ScaSML n=rho=4, GP 4000+800 collocation points (phi=16800); K-streamed tcgen05 kernel"),
    "C4": dict(d=60, nd=1000, nb=200, n=4, rho=None, variant="full_history", M=3,
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
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_oracle_rate(cfg, dom, bdy, X, alpha, idx, quadrature, budget_s):
    """Time the CPU oracle (NumPy + OpenBLAS, all host threads) on a bounded sample of the same workload."""
    from oracle.equation import EquationOracle
    from oracle.gp import GPOracle
    from oracle.solvers import ScaSMLFullHistoryOracle, ScaSMLOracle
    eq_o = EquationOracle(cfg["d"] + 1)
    gp_o = GPOracle(eq_o, idx_set=idx)
    gp_o.x_t_domain, gp_o.x_t_boundary = dom, bdy
    gp_o.N_domain, gp_o.N_boundary = len(dom), len(bdy)
    gp_o.right_vector = alpha.reshape(-1, 1)
    fh = cfg["variant"] == "full_history"

    def run(nb):
        s = (ScaSMLFullHistoryOracle if fh else ScaSMLOracle)(eq_o, gp_o, cast=False, true_gl=(quadrature != "reference"))
        t0 = time.perf_counter()
        if fh:
            s.u_solve(cfg["n"], None, X[:nb], M=cfg["M"])
        else:
            s.u_solve(cfg["n"], cfg["rho"], X[:nb])
        dt = time.perf_counter() - t0
        return dt, s
    dt, s = run(1)
    nb = int(max(1, min(len(X), budget_s / max(dt, 1e-3))))
    if nb > 1:
        dt, s = run(nb)
    return nb, dt, s


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
    config = {"workload": cfg["desc"], "config": args.config, "test_points_per_gpu": args.batch, "test_points_total": B_total,
              "parallelism": f"sample-sharded x{world}" if world > 1 else "single GPU",
              "quadrature_tables": args.quadrature, "l2": "level point buffers (GBs) exceed L2 every step",
              "sample_point_unit": "executed sample points (level-0 terminals the reference draws and discards are elided)"}

    # ------------------------------------------------------------------ reference arm: CPU oracle ----------
    if args.impl == "reference":
        if rank != 0:
            return
        # fit on CPU (small sample of the path is what is timed; the fit is setup)
        from oracle.equation import EquationOracle
        from oracle.gp import GPOracle
        eq_o = EquationOracle(d + 1)
        gp_o = GPOracle(eq_o, idx_set=idx)
        t0 = time.perf_counter()
        gp_o.GPsolver(dom, bdy)
        fit_s = time.perf_counter() - t0
        alpha = gp_o.right_vector[:, 0]
        rates = []
        nb = dt = None
        for it in range(args.warmup + args.steps):
            nb, dt, s = cpu_oracle_rate(cfg, dom, bdy, X, alpha, idx, args.quadrature, args.cpu_seconds / max(args.steps, 1))
            if it >= args.warmup:
                rates.append(s.sample_points_executed / dt if hasattr(s, "sample_points_executed") else _exec_sp(s, nb) / dt)
        val = float(np.mean(rates))
        line = {"impl": "reference", "metric": "ScaSML correction sample-points/s", "value": val, "unit": "sample-points/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "sample-points/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{nb} test points of the same workload per step (NumPy+OpenBLAS oracle, all host threads)",
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
    gp.route = _lib.ROUTE_TC if route == "tc" else _lib.ROUTE_F64
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gp.GPsolver(dom, bdy)
    torch.cuda.synchronize()
    fit_ms = 1e3 * (time.perf_counter() - t0)

    solver = (ScaSML_full_history if fh else ScaSML)(eq, gp)
    solver.route = gp.route
    solver.quadrature = args.quadrature
    solver.distributed = world > 1
    x_dev = _lib.to_device(X)
    torch.cuda.synchronize()

    def step_device():
        uz = solver._uz_device(n, rho, x_dev, M)
        (uh,) = gp._eval(x_dev, _lib.EVAL_U)
        return uz[:, 0] + uh

    def step_e2e():
        return solver.u_solve(n, rho, X, M) if fh else solver.u_solve(n, rho, X)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step_device()
    # device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
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
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        ex = torch.tensor([float(stats["executed_points"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(ex, op=dist.ReduceOp.SUM)
        executed_total = float(ex[0])
    else:
        executed_total = float(stats["executed_points"])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

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
    eval_s = tstats["eval_time_ns"] * 1e-9
    achieved_tf = tstats["eval_flops"] / max(eval_s, 1e-12) / 1e12
    if route == "tc":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_note = "measured cuBLAS bf16 sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PFLOP/s sustained"
        flop_note = ("kind::f16 tcgen05, two chained GEMMs, operands split hi+lo in f16 (2 distance passes, 3 coefficient products); "
                     "credited flops are the algorithmic 2(d+1) per pair-distance only")
    else:
        # FP64 route: the governing pipe is the FP64 FMA pipe; measure its peak here with a torch fp64 GEMM
        a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        torch.matmul(a, a)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            torch.matmul(a, a)
        g1.record()
        torch.cuda.synchronize()
        peak = 3 * 2 * 4096 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
        peak_note = "FP64 route: cuBLAS fp64 4096^3 GEMM measured in this run (MEASURED_PEAKS.json has no fp64 figure)"
        flop_note = "FP64 FMA pipe (SIMT contraction); credited flops are the algorithmic 2(d+1) per pair-distance"
    # DRAM traffic of the evaluation launches: ncu --set full on one launch (profiles/r1_ncu_eval_pde.md) measured
    # dram__bytes_read + write = 827 B per evaluated point on the PDE launch (algorithmic: 808 B point row + 8..32 B of outputs)
    traffic = 827.0 * tstats["eval_points_total"] / max(tstats["eval_launches"], 1) if route == "tc" else None
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak,
                "traffic": traffic, "traffic_note": "bytes per evaluation launch (mean): 827 B/point from the ncu capture profiles/r1_ncu_eval_pde.md x points per launch",
                "kernel": "fused surrogate evaluation (gp_eval*.cu)", "peak_source": peak_note,
                "note": flop_note, "eval_share_of_step": tstats["eval_time_ns"] / max(
                    tstats["eval_time_ns"] + tstats["sample_time_ns"] + tstats["reduce_time_ns"], 1),
                "eval_ms": 1e-6 * tstats["eval_time_ns"], "sample_ms": 1e-6 * tstats["sample_time_ns"],
                "reduce_ms": 1e-6 * tstats["reduce_time_ns"], "eval_launches": tstats["eval_launches"]}

    # HBM-bound kernels of the step (samplers write, reduction reads the level point buffers: (d+1) * 8 B per sample point)
    hbm_peak = peaks.get("hbm_gbs", 6554.0) if peaks else 6554.0
    pt_bytes = (d + 1) * 8.0 * float(stats["executed_points"])
    hbm = {"unit": "GB/s", "peak": hbm_peak, "bytes_per_point": (d + 1) * 8,
           "sampler": {"achieved": pt_bytes / max(tstats["sample_time_ns"], 1), "frac": pt_bytes / max(tstats["sample_time_ns"], 1) / hbm_peak},
           "reduction": {"achieved": pt_bytes / max(tstats["reduce_time_ns"], 1), "frac": pt_bytes / max(tstats["reduce_time_ns"], 1) / hbm_peak}}
    roofline["hbm_kernels"] = hbm

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        nb, dt, s = cpu_oracle_rate(cfg, dom, bdy, X, gp.right_vector[:, 0], idx, args.quadrature, args.cpu_seconds)
        cpu = {"value": _exec_sp(s, nb) / dt, "unit": "sample-points/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{nb} test points of the same workload, {dt:.1f} s (NumPy+OpenBLAS oracle, all host threads)"}

    D = d + 1
    line = {"metric": "ScaSML correction sample-points/s", "value": value, "unit": "sample-points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if route == "f64" else "f16 (split operands, fp32 tensor-core accumulate, f64 final contraction)",
            "data": "synthetic", "config": dict(config, route=route),
            "reference_equivalent_value": ref_equiv, "fit_ms": fit_ms, "newton_steps": gp.newton_steps,
            "finite_fraction": finite,
            "e2e": {"value": e2e_value, "unit": "sample-points/s", "h2d_bytes_per_step": B_total * D * 8,
                    "d2h_bytes_per_step": B_total * D * 8 + B_total * 8, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int((stats["launches"] + 1) * args.steps),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary()}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def _exec_sp(solver, nb):
    """Executed sample points of an oracle run: everything drawn minus the discarded level-0 terminals."""
    # the oracle counts level-0 terminals in sample_points but does not draw them (see oracle/solvers.py n == 0)
    return float(getattr(solver, "sample_points_executed", solver.sample_points))


if __name__ == "__main__":
    main()
