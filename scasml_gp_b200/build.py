"""Build the C-ABI libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU).

  libscasml_b200.so      the product: include/scasml_b200.h, nothing else exported
  libscasml_b200_dbg.so  the same sources + test hooks / micro-benchmarks (include/scasml_b200_debug.h), compiled with
                         -DSCASML_DEBUG_HOOKS (timeline stamps and experiment flags inside the tcgen05 kernel);
                         loaded only by tests/ and tools/
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscasml_b200.so")
LIB_DBG = os.path.join(HERE, "libscasml_b200_dbg.so")
SOURCES = ["abi.cu", "comm.cu", "gp_eval.cu", "gp_eval_tc.cu", "gp_fit.cu", "picard.cu"]
# debug library: objects shared with the product except the files that look at SCASML_DEBUG_HOOKS
DBG_ONLY = ["abi_debug.cu", "tc_bench.cu"]
DBG_REBUILT = ["gp_eval_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Wno-deprecated-gpu-targets"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not (os.path.exists(LIB) and os.path.exists(LIB_DBG)):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_DBG))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", h) for h in ("scasml_b200.h", "scasml_b200_debug.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    jobs = [(src, [], os.path.join(bdir, src.replace(".cu", ".o"))) for src in SOURCES]
    jobs += [(src, ["-DSCASML_DEBUG_HOOKS"], os.path.join(bdir, src.replace(".cu", "_dbg.o"))) for src in DBG_REBUILT + DBG_ONLY]
    procs = []
    for src, extra, obj in jobs:
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
    obj = {(src, bool(extra)): o for src, extra, o in jobs}
    prod = [obj[(s, False)] for s in SOURCES]
    dbg = [obj[(s, True)] if s in DBG_REBUILT else obj[(s, False)] for s in SOURCES] + [obj[(s, True)] for s in DBG_ONLY]
    subprocess.run([nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB, *prod, "-lcudart", "-ldl"], check=True)
    subprocess.run([nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB_DBG, *dbg, "-lcudart", "-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
