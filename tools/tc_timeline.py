"""Print the SM-clock timeline of one CTA of the tcgen05 evaluation kernel (debug aid; needs a GPU)."""
import sys
import numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear

d, nd, nb = 100, 1000, 200
dom, bdy, X = gen_points(d, nd, nb, 148 * 128 * 4)
eq = Grad_Dependent_Nonlinear(d + 1)
gp = GP_Grad_Dependent_Nonlinear(eq)
gp._bind(dom, bdy)
gp.set_right_vector(np.random.default_rng(0).standard_normal(4 * nd + nb) * 0.1)
xd = _lib.to_device(X)
R = xd.shape[0]
lib = _lib.load_debug()          # stamps / experiment flags exist only in the debug build
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0     # 1 skip stage-2 MMAs, 2 skip epilogue arithmetic, 4 skip low-half pass
for mode, name in ((_lib.EVAL_U, "U"), (_lib.EVAL_UG, "UG"), (_lib.EVAL_PDE, "PDE")):
    stamps = torch.zeros(1024, dtype=torch.int64, device="cuda")
    scratch = torch.empty(4 * R, dtype=torch.float64, device="cuda")
    for rep in range(2):
        _lib.check(lib.scasml_debug_tc_timeline(gp._handle, _lib.ptr(xd), R, mode, 100 | (flags << 24), _lib.ptr(stamps), _lib.ptr(scratch), _lib.stream_ptr()))
    torch.cuda.synchronize()
    t = stamps.cpu().numpy().astype(np.int64)
    t0 = t[0]
    nitem = int(((t[4:244] > 0).sum()) // 4)
    starts = [t[4 + 4 * w] for w in range(nitem)]
    period = (starts[-3] - starts[3]) / max(1, nitem - 6) if nitem > 8 else 0
    print(f"== {name} flags={flags}: steady-state period {period:.0f} cycles/item")
    print(f"== {name}: entry 0, scatter {t[1]-t0}, prologue {t[2]-t0}, exit {t[3]-t0} cycles; items {nitem}")
    f = [int(x - t0) for x in t[244:249]]
    print(f"   tile boundary (stamps of the CTA's SECOND point tile): tile-0 t_full committed {f[4]}, stage-1 issuer enters tile 1 {f[0]}, "
          f"A images copied {t[2]-t0}, contraction warp 0 saw t_full(0) {f[2]}, tile-0 outputs written {f[3]}")
    for w in range(min(nitem, 44)):
        a, b, c, e = (t[4 + 4 * w + i] - t0 for i in range(4))
        f = [int(t[256 + 8 * w + i] - t0) for i in range(7)]
        print(f"  pair {w:2d}: S1 issuer reaches {a:7d} | slot free {f[0]:7d} rows landed {f[1]:7d} issued {f[2]:7d} committed {f[3]:7d} | epilogue {c:7d}..{e:7d} ({e-c:5d}) | "
              f"S2 issuer: P ready {f[4]:7d} images landed {f[5]:7d} issued {f[6]:7d} committed {b:7d}")
