"""Stress the Picard solvers on the tcgen05 route at full batch size: repeated solves (fresh keys each), NaN counts, and at d <= 20 the
FP64 route beside it.   python tools/stress_tc.py d n B reps   (needs a GPU)"""
import sys, numpy as np, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
from scasml_gp_b200.solvers.ScaSML import ScaSML
d = int(sys.argv[1]); n = int(sys.argv[2]); B = int(sys.argv[3]); reps = int(sys.argv[4])
dom, bdy, X = gen_points(d, 1000, 200, B)
eq = Grad_Dependent_Nonlinear(d + 1)
gp = GP_Grad_Dependent_Nonlinear(eq)
gp._bind(dom, bdy)
gp.set_right_vector(np.random.default_rng(0).standard_normal(4 * 1000 + 200) * 0.01)
res = {}
for route in (_lib.ROUTE_TC, _lib.ROUTE_F64) if d <= 20 else (_lib.ROUTE_TC,):
    s = ScaSML(eq, gp); s.route = route; s.quadrature = "gauss_legendre"
    outs = []
    for r in range(reps):
        try:
            s.u_solve(n, n, X)
        except Exception as e:
            print("route", route, "rep", r, "FAILED", str(e)[:200]); sys.exit(1)
        outs.append(s.last_raw.copy())
        print("route", route, "rep", r, "nan", int(np.isnan(outs[-1]).sum()), "absmax", float(np.nanmax(np.abs(outs[-1]))), flush=True)
    res[route] = outs
if len(res) == 2:
    for r in range(reps):
        a, b = res[_lib.ROUTE_TC][r], res[_lib.ROUTE_F64][r]
        print("rep", r, "max diff tc vs f64", float(np.nanmax(np.abs(a - b))), "nan mismatch", int((np.isnan(a) != np.isnan(b)).sum()))
