"""tcgen05.mma micro-benchmark (needs a GPU): cycles per instruction vs N, independent chains, A source."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scasml_gp_b200 import _lib as lib
L = lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 512
print("N chains mode  issue/instr  total/instr")
for ts in (0, 1, 2):
    for N in (64, 128, 256):
        for ch in (1, 2, 3, 4):
            if ch * N > 448:
                continue
            for rep in range(2):
                lib.check(L.scasml_debug_tc_mma_bench(N, ch, ts, iters, lib.ptr(out), lib.stream_ptr()))
            torch.cuda.synchronize()
            a, b = out.cpu().tolist()
            print(f"{N:4d} {ch} {["SS","TS","TS4"][ts]}   {a/iters:8.1f}   {b/iters:8.1f}")
