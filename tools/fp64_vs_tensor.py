"""Is the FP64 SIMT pipe slowed down while the tensor pipe is busy?  (debug aid; needs a GPU)
16 warps per SM run a DFMA (FFMA) loop of 8 independent chains, with and without the MMA warp issuing TS-mode N = 128 MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scasml_gp_b200 import _lib as lib
L = lib.load_debug()
out = torch.zeros(8, dtype=torch.int64, device="cuda")
iters = 448 * 4
for mode, name in ((13, "DFMA alone"), (12, "DFMA + MMAs"), (15, "FFMA alone"), (14, "FFMA + MMAs")):
    for _ in range(2):
        lib.check(L.scasml_debug_tc_pipe_bench(mode, 128, iters, lib.ptr(out), lib.stream_ptr()), L)
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"{name:12s}: {o[0] / (iters * 8):6.2f} cycles per instruction per warp (16 warps/SM)" + (f"; MMA {o[2] / max(o[3], 1):.1f} cyc/instr" if mode in (12, 14) else ""))
