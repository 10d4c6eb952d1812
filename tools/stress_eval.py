"""Stress the tcgen05 evaluation kernel: many point tiles per CTA, repeated launches; reports NaNs and run-to-run differences (the kernel
is deterministic: any difference is a race).  Needs a GPU.   python tools/stress_eval.py [d] [tiles per SM] [reps] [debug flags] [nd] [nb]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear

d = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 40
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
flags = int(sys.argv[4]) if len(sys.argv) > 4 else -1     # >= 0: debug library with these experiment flags (gp_eval_tc.cu::TcDev::dbg_flags; 4 = watchdog)
nd = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
nb = int(sys.argv[6]) if len(sys.argv) > 6 else 200
R = 148 * 128 * tiles + 77
dom, bdy, X = gen_points(d, nd, nb, R)
eq = Grad_Dependent_Nonlinear(d + 1)
gp = GP_Grad_Dependent_Nonlinear(eq)
gp._bind(dom, bdy)
gp.set_right_vector(np.random.default_rng(0).standard_normal(4 * nd + nb) * 0.01)
gp.route = _lib.ROUTE_TC
xd = _lib.to_device(X)
bad = 0
NOUT = {_lib.EVAL_U: 1, _lib.EVAL_UG: 2, _lib.EVAL_PDE: 4}
for mode, name in ((_lib.EVAL_U, "U"), (_lib.EVAL_UG, "UG"), (_lib.EVAL_PDE, "PDE")):
    first = None
    for r in range(reps):
        if flags < 0:
            outs = [o.clone() for o in gp._eval(xd, mode, nout=NOUT[mode])]
        else:
            stamps = torch.zeros(1024 + 16 * 148, dtype=torch.int64).pin_memory()     # host memory: survives a trap
            scratch = torch.zeros(4 * R, dtype=torch.float64, device="cuda")
            _lib.check(_lib.load_debug().scasml_debug_tc_timeline(gp._handle, _lib.ptr(xd), R, mode, 100 | (flags << 24), _lib.ptr(stamps),
                                                                  _lib.ptr(scratch), _lib.stream_ptr()))
            outs = [scratch[i * R:(i + 1) * R].clone() for i in range(NOUT[mode])]
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("launch failed:", str(e).splitlines()[0], "| watchdog (line, CTA, warp, parity, barrier):", stamps[960:965].tolist() if flags >= 0 else None, flush=True)
            if flags >= 0:
                for w, v in enumerate(stamps[900:923].tolist()):
                    if v:
                        print(f"    warp {w}: stuck at source line {v & 0xffff}, parity {(v >> 16) & 1}, CTA {v >> 32}")
                for cta in sorted({int(stamps[961]), *[int(v) >> 32 for v in stamps[900:923].tolist() if v]}):
                    pr = stamps[1024 + 16 * cta:1024 + 16 * cta + 9].tolist()
                    names = ["producer (b1 fills, pair)", "-", "S1A", "S1B", "S2A", "S2B", "contraction (tile, class)", "epilogue group 0", "epilogue group 1"]
                    print(f"    CTA {cta} progress (tile, pair): " + ", ".join(f"{n} {v >> 16}/{v & 0xffff}" for n, v in zip(names, pr) if n != "-"))
            os._exit(1)
        nan = sum(int(torch.isnan(o).sum()) for o in outs)
        diff = 0 if first is None else sum(int((o != f).sum()) for o, f in zip(outs, first))
        if first is None:
            first = outs
        bad += nan + diff
        print(f"{name} rep {r}: nan {nan}, entries differing from rep 0: {diff}", flush=True)
        if diff:
            for k, (o, f) in enumerate(zip(outs, first)):
                idx = torch.nonzero(o != f).flatten()
                if idx.numel():
                    rel = ((o[idx] - f[idx]).abs() / f[idx].abs().clamp_min(1e-300)).max().item()
                    rows = idx.tolist()
                    print(f"    out{k}: {idx.numel()} rows, first {rows[:6]} last {rows[-3:]}, tiles {sorted(set(i // 128 for i in rows))[:8]}, lanes mod 32 {sorted(set(i % 32 for i in rows))[:8]}, max rel diff {rel:.3e}")
print("BAD" if bad else "CLEAN")
