"""tcgen05.mma micro-benchmark (needs a GPU): cycles per instruction vs N, independent chains, A source, and
concurrent tcgen05.ld traffic from the other warps of the CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scasml_gp_b200 import _lib as lib
L = lib.load_debug()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
iters = 512
names = {0: "SS loop", 1: "TS loop", 2: "TS unrolled", 4: "TS unrolled + light LDTM", 8: "TS unrolled + heavy LDTM"}
print("N chains mode  issue/instr  total/instr")
for ts in (0, 1, 2, 4, 8):
    for N in (64, 128, 256):
        for ch in (1, 2):
            if ch * N > 256:
                continue
            for rep in range(2):
                lib.check(L.scasml_debug_tc_mma_bench(N, ch, ts, iters, lib.ptr(out), lib.stream_ptr()))
            torch.cuda.synchronize()
            a, b = out.cpu().tolist()[:2]
            print(f"{N:4d} {ch} {names[ts]:26s} {a/iters:8.1f}   {b/iters:8.1f}")
