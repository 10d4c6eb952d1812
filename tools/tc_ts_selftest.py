import sys, numpy as np
sys.path.insert(0, "/root/repo")
import torch
from scasml_gp_b200 import _lib as lib
for (K, N) in [(64, 64), (128, 64), (128, 32)]:
    rng = np.random.default_rng(K + N)
    A = rng.standard_normal((128, K)).astype(np.float16); B = rng.standard_normal((N, K)).astype(np.float16)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Dd = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    lib.check(lib.load_debug().scasml_debug_tc_gemm(lib.ptr(Ad), lib.ptr(Bd), lib.ptr(Dd), K, N, 1, 64, 100, 32, lib.stream_ptr()))
    torch.cuda.synchronize()
    got = Dd.cpu().numpy()
    print("TS", K, N, "max err", float(np.nanmax(np.abs(got - want))), "nan", int(np.isnan(got).sum()))
