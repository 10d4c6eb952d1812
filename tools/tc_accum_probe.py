"""How does tcgen05.mma (kind::f16, FP32 accumulators in tensor memory) round its accumulation?  (debug aid; needs a GPU)

D[128 x 64] = A[128 x K] B[64 x K]^T through the plumbing self-test kernel (one accumulator chain of K / 16 MMAs), compared with the
exact float64 product of the same float16 operands and with three models of the accumulation: exact sum rounded once, and
per-MMA round-to-nearest / round-toward-zero to FP32 (exact inside an MMA).  Errors in ulps of the final value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scasml_gp_b200 import _lib

dbg = _lib.load_debug()


def gemm(A, B):
    K, N = A.shape[1], B.shape[0]
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Dd = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(dbg.scasml_debug_tc_gemm(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(Dd), K, N, 1, 64, 2, 32, _lib.stream_ptr()), dbg)
    torch.cuda.synchronize()
    return Dd.cpu().numpy().astype(np.float64)


def rz32(x):
    f = x.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float64)


def model(A, B, mode):
    acc = np.zeros((A.shape[0], B.shape[0]))
    for k0 in range(0, A.shape[1], 16):
        acc = acc + A[:, k0:k0 + 16].astype(np.float64) @ B[:, k0:k0 + 16].astype(np.float64).T
        acc = rz32(acc) if mode == "rz" else acc.astype(np.float32).astype(np.float64)
    return acc


rng = np.random.default_rng(0)
for name, gen in (("positive terms", lambda s: rng.uniform(0.5, 1.0, s)), ("mixed signs", lambda s: rng.standard_normal(s))):
    for K in (64, 256):
        A, B = gen((128, K)).astype(np.float16), gen((64, K)).astype(np.float16)
        exact = A.astype(np.float64) @ B.astype(np.float64).T
        got = gemm(A, B)
        ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
        line = f"{name:15s} K={K:4d} ({K // 16:2d} MMAs)  |exact| ~ {np.abs(exact).mean():8.3f}"
        for lab, val in (("tensor core", got), ("model RN/MMA", model(A, B, 'rn')), ("model RZ/MMA", model(A, B, 'rz'))):
            e = (val - exact) / ulp
            line += f" | {lab}: mean {np.mean(e * np.sign(exact)):+7.3f} rms {np.sqrt(np.mean(e * e)):6.3f} ulp"
        print(line)
        print(f"{'':15s} bit-equal to model RZ/MMA: {np.mean(got == model(A, B, 'rz')):.3f}, to RN/MMA: {np.mean(got == model(A, B, 'rn')):.3f}")
