"""Where the wall time of GP.GPsolver goes (host phases, steady state): bind / fit call / alpha copy / predict (debug aid; needs a GPU)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear

d, nd, nb = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (100, 1000, 200)
dom, bdy, X = gen_points(d, nd, nb, 64)
eq = Grad_Dependent_Nonlinear(d + 1)
gp = GP_Grad_Dependent_Nonlinear(eq, idx_set=np.random.default_rng(0).choice(d, 5, replace=False))
lib = _lib.load()


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"   {label:28s} {1e3 * (t1 - t0):8.2f} ms")
    return time.perf_counter()


for rep in range(3):
    print(f"fit {rep}")
    torch.cuda.synchronize()
    t00 = t0 = time.perf_counter()
    gp._bind(dom, bdy)
    t0 = tick("bind (handle, centres)", t0)
    N = gp.N_domain
    g_bdy = _lib.to_device(np.asarray(gp.bdy_g(gp.x_t_boundary), dtype=np.float64))
    sol0_d = _lib.to_device(np.random.default_rng(0).standard_normal(3 * N) * 1e-3)
    ws_bytes = lib.scasml_gp_fit_workspace_bytes(gp._handle)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    sol_out = torch.empty(3 * N, dtype=torch.float64, device="cuda")
    hist = (C.c_double * 21)()
    steps = C.c_int(0)
    t0 = tick("inputs + workspace", t0)
    _lib.check(lib.scasml_gp_fit(gp._handle, _lib.ptr(g_bdy), _lib.ptr(sol0_d), 20, 1e-4, 1e-5, 1, _lib.ptr(ws), ws_bytes,
                                 _lib.ptr(sol_out), hist, C.byref(steps), _lib.stream_ptr()))
    t0 = tick(f"scasml_gp_fit ({steps.value} steps)", t0)
    alpha = torch.empty(gp.phi_dim, dtype=torch.float64, device="cuda")
    _lib.check(lib.scasml_gp_get_alpha(gp._handle, _lib.ptr(alpha), _lib.stream_ptr()))
    gp.right_vector = alpha.cpu().numpy()[:, None]
    t0 = tick("alpha to host", t0)
    for route in (_lib.ROUTE_F64, _lib.ROUTE_TC):
        gp.route = route
        gp.predict(dom)
        t0 = tick(f"predict(x_dom) route {route}", t0)
    print(f"   total {1e3 * (time.perf_counter() - t00):8.2f} ms")
