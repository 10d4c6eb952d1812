"""Small invocations of every tcgen05 kernel variant for compute-sanitizer (tools/sanitize.sh): the three evaluation classes
of the resident-operand kernel at d = 20 (2 k-steps) and d = 100 (7 k-steps, two K blocks), the K-streamed kernel at d = 300,
and one level-2 solve through the samplers / reductions.  Sizes are a few point tiles: the sanitizer slows kernels 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
from scasml_gp_b200.solvers.ScaSML import ScaSML

which = sys.argv[1] if len(sys.argv) > 1 else "all"
for d, nd, nb, R in ((20, 130, 40, 128 * 2 + 37), (100, 130, 70, 128 * 2 + 37), (300, 70, 20, 128 + 9)):
    if which not in ("all", str(d)):
        continue
    dom, bdy, X = gen_points(d, nd, nb, R)
    eq = Grad_Dependent_Nonlinear(d + 1)
    gp = GP_Grad_Dependent_Nonlinear(eq)
    gp._bind(dom, bdy)
    gp.set_right_vector(np.random.default_rng(0).standard_normal(4 * nd + nb) * 0.1)
    gp.route = _lib.ROUTE_TC
    xd = _lib.to_device(X)
    for mode, nout in ((_lib.EVAL_U, 1), (_lib.EVAL_TERMINAL, 1), (_lib.EVAL_UG, 2), (_lib.EVAL_PDE, 4)):
        outs = gp._eval(xd, mode, nout)
        torch.cuda.synchronize()
        assert all(bool(torch.isfinite(o).all()) for o in outs), (d, mode)
    if d <= 100:
        s = ScaSML(eq, gp)
        s.route = _lib.ROUTE_TC
        u = s.u_solve(2, 2, X[:9])
        assert np.all(np.isfinite(u.astype(np.float64)))
    print(f"sanitize target d={d}: ok", flush=True)
