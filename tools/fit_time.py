"""Wall time of repeated GP fits at the C3 size (d = 100, 1000 + 200 collocation points), steady state (debug aid; needs a GPU)."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import gen_points
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
for d, nd, nb in ((100, 1000, 200),):
    dom, bdy, X = gen_points(d, nd, nb, 64)
    eq = Grad_Dependent_Nonlinear(d + 1)
    for rep in range(3):
        gp = GP_Grad_Dependent_Nonlinear(eq, idx_set=np.random.default_rng(0).choice(d, 5, replace=False))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gp.GPsolver(dom, bdy)
        torch.cuda.synchronize(); print(d, "fit", rep, round(1e3 * (time.perf_counter() - t0), 1), "ms", gp.newton_steps, "steps", float(np.abs(gp.right_vector).sum()))
