"""Time the pivoted-LU solve of the fit's fallback path on random n x n systems (debug aid; needs a GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scasml_gp_b200 import _lib
lib = _lib.load()
for n in ([int(a) for a in sys.argv[1:]] or [3000, 6000, 12000]):
    g = torch.Generator(device="cuda").manual_seed(0)
    H = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    for rep in range(1 if len(sys.argv) > 1 else 2):
        Hd, bd = H.clone(), b.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(_lib.load_debug().scasml_debug_lu_solve(_lib.ptr(Hd), n, _lib.ptr(bd), _lib.stream_ptr()))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r = (H @ bd - b).norm() / b.norm()
    print(f"n={n}: {1e3*dt:.1f} ms, {2*n**3/3/dt/1e12:.2f} TFLOP/s, residual {float(r):.2e}")
