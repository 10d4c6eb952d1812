"""torchrun --nproc-per-node N tools/nrank_check.py: the N-rank sharded solve (NCCL all-reduce of the weighted partial
sums) against the single-GPU solve of the same test points, on the hardware.  Prints NRANK_CHECK_OK."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import gen_points
from scasml_gp_b200 import _lib
from scasml_gp_b200.equations.equations import Grad_Dependent_Nonlinear
from scasml_gp_b200.models.GP import GP_Grad_Dependent_Nonlinear
from scasml_gp_b200.solvers.ScaSML import ScaSML
from scasml_gp_b200.solvers.ScaSML_full_history import ScaSML_full_history


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, nd, nb, B = 20, 200, 40, 37
    dom, bdy, X = gen_points(d, nd, nb, B)
    eq = Grad_Dependent_Nonlinear(d + 1)
    gp = GP_Grad_Dependent_Nonlinear(eq)
    gp.GPsolver(dom, bdy)
    worst = 0.0
    for cls, args in ((ScaSML, (3, 3, X)), (ScaSML_full_history, (3, None, X, 3))):
        for route in (_lib.ROUTE_F64, _lib.ROUTE_TC):
            one = cls(eq, gp)
            one.route = route
            one.u_solve(*args)
            many = cls(eq, gp)
            many.route = route
            many.distributed = True
            many.u_solve(*args)
            diff = float(np.nanmax(np.abs(many.last_raw - one.last_raw)))
            assert np.array_equal(np.isnan(many.last_raw), np.isnan(one.last_raw))
            assert diff < 1e-12, (cls.__name__, route, diff)          # same terms, summed in a different order
            assert many.evaluation_counter == one.evaluation_counter
            worst = max(worst, diff)
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"NRANK_CHECK_OK world={world} max_abs_diff={float(t[0]):.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
