"""Accuracy of the tcgen05 evaluation route against the FP64 route on a fitted C3-sized surrogate (debug aid; needs a GPU):
rms and max absolute difference of u_hat, sum_i d_i u_hat and the PDE residual, relative to the FP64 values' scale."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear

d, nd, nb = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (100, 1000, 200)
dom, bdy, X = gen_points(d, nd, nb, 4096)
X = X + 0.01 * np.random.default_rng(1).standard_normal(X.shape)     # not float16-representable
eq = Grad_Dependent_Nonlinear(d + 1)
gp = GP_Grad_Dependent_Nonlinear(eq, idx_set=np.random.default_rng(0).choice(d, 5, replace=False))
gp.GPsolver(dom, bdy, GN_steps=20 if d <= 200 else 3)
res = {}
for route in (_lib.ROUTE_F64, _lib.ROUTE_TC):
    gp.route = route
    u = gp.predict_raw(X)
    eps, dv, lp, dt = gp.pde_terms_raw(X)
    uu, G = [t.cpu().numpy() for t in gp._eval(X, _lib.EVAL_UG, nout=2)]
    res[route] = dict(u=u, G=G, eps=eps, dv=dv, lp=lp, dt=dt)
a, b = res[_lib.ROUTE_F64], res[_lib.ROUTE_TC]
for k in a:
    s = max(1e-300, float(np.abs(a[k]).max()))
    print(f"{k:4s} scale {s:9.3e}  rms diff / scale {np.sqrt(np.mean((a[k]-b[k])**2))/s:9.2e}  max diff / scale {np.max(np.abs(a[k]-b[k]))/s:9.2e}"
          f"  mean diff / scale {np.mean(b[k]-a[k])/s:10.2e}  mean relative diff {np.mean((b[k]-a[k])/np.where(a[k] == 0, 1, a[k])):10.2e}")
