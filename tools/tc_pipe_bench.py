"""Epilogue-pipe micro-benchmarks (needs a GPU): TMEM load/store bandwidth, MUFU rate, the ex2 + f16 hi/lo split chunk of
the evaluation epilogue, tcgen05.mma cost vs N (SS / TS), and MMA <-> epilogue interference.  See csrc/tc_bench.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scasml_gp_b200 import _lib as lib
L = lib.load_debug()
out = torch.zeros(8, dtype=torch.int64, device="cuda")
iters = 448 * 2


def run(mode, N=64):
    for _ in range(2):
        lib.check(L.scasml_debug_tc_pipe_bench(mode, N, iters, lib.ptr(out), lib.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().tolist()


o = run(0)
print(f"LDTM x16: {o[0]/iters:.1f} cyc per 4 loads/warp -> {16*4*2048*iters/o[0]:.1f} B/clk/SM")
o = run(1)
print(f"STTM x8 : {o[0]/iters:.1f} cyc per 8 stores/warp -> {16*8*1024*iters/o[0]:.1f} B/clk/SM")
o = run(2)
print(f"MUFU ex2: {o[0]/iters:.1f} cyc per 16 ex2/thread -> {512*16*iters/o[0]:.2f} ex2/clk/SM")
o = run(3)
print(f"split chunk (ld16, 16 ex2, hi/lo, 2 st8): {o[0]/iters:.1f} cyc per chunk/warp -> {512*16*iters/o[0]:.2f} pairs/clk/SM")
for N in (64, 128, 192, 256):
    o = run(4, N)
    print(f"split chunk + SS MMA N={N}: {o[0]/iters:.1f} cyc per chunk/warp ({512*16*iters/o[0]:.2f} pairs/clk/SM); "
          f"MMA {o[2]/max(o[3],1):.1f} cyc/instr")
for mode, name in ((5, "SS"), (6, "TS")):
    for N in (16, 32, 48, 64, 96, 128, 192):
        o = run(mode, N)
        print(f"MMA {name} N={N:3d}: issue {o[1]/o[3]:.1f}  total {o[2]/o[3]:.1f} cyc/instr")
for mode, name in ((7, "TS + commit/7"), (8, "TS + commit/7 + poll")):
    for N in (16, 64, 128):
        o = run(mode, N)
        print(f"MMA {name} N={N:3d}: issue {o[1]/o[3]:.1f}  total {o[2]/o[3]:.1f} cyc/instr")
for N in (16, 64, 128):
    o = run(9, N)
    print(f"split chunk + TS MMA N={N}: {o[0]/iters:.1f} cyc per chunk/warp ({512*16*iters/o[0]:.2f} pairs/clk/SM); MMA {o[2]/max(o[3],1):.1f} cyc/instr")
for mode, name in ((10, "TS interleaved 1 x N=128 + 2 x N=16"), (11, "TS batched 14 x N=128 then 24 x N=16")):
    o = run(mode, 128)
    print(f"MMA {name}: {38 * o[2] / o[3]:.0f} cycles per (14 + 24) MMAs (expected 14 x 68 + 24 x 17 = 1360)")
