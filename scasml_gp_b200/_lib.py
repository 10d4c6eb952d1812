"""ctypes binding of libscasml_b200.so (include/scasml_b200.h) + device-buffer helpers.

PyTorch is used only to own device memory, pinned host staging buffers and the current stream.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscasml_b200.so")
LIB_DBG_PATH = os.path.join(HERE, "libscasml_b200_dbg.so")      # test hooks + micro-benchmarks (include/scasml_b200_debug.h)

MAX_LEVEL = 8
MAX_Q = 8
EVAL_U, EVAL_TERMINAL, EVAL_UG, EVAL_PDE = 0, 1, 2, 3
ROUTE_F64, ROUTE_TC = 0, 1
ERR_NUMERIC = 3


class PicardParams(C.Structure):
    _fields_ = [
        ("variant", C.c_int), ("scasml", C.c_int), ("n", C.c_int), ("d", C.c_int), ("M", C.c_int), ("qmax", C.c_int),
        ("Qrow", C.c_int * MAX_LEVEL), ("Mfrow", C.c_int * MAX_LEVEL), ("Mgrow", C.c_int * (MAX_LEVEL + 1)),
        ("c", C.c_double * (MAX_Q * MAX_Q)), ("w", C.c_double * (MAX_Q * MAX_Q)),
        ("T", C.c_double), ("mu", C.c_double), ("sigma", C.c_double), ("clip", C.c_double),
        ("stale_delta", C.c_int), ("cast_levels", C.c_int), ("seed", C.c_uint), ("key_counter", C.c_uint),
        ("rank", C.c_int), ("world", C.c_int), ("gid0", C.c_longlong), ("timing", C.c_int), ("reserved", C.c_int),
    ]


class PicardStats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("keys_used", "eval_counter", "sample_points", "executed_points",
                                            "n_calls", "launches", "eval_points_total", "eval_launches",
                                            "eval_time_ns", "sample_time_ns", "reduce_time_ns", "eval_flops")]


_SIGNATURES = {
    "scasml_last_error": (C.c_char_p, []),
    "scasml_abi_version": (C.c_int, []),
    "scasml_set_normal_table": (C.c_int, [C.c_void_p]),
    "scasml_comm_unique_id": (C.c_int, [C.c_void_p]),
    "scasml_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "scasml_allreduce_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "scasml_comm_destroy": (C.c_int, [C.c_void_p]),
    "scasml_geometry_points": (C.c_int, [C.c_uint, C.c_uint, C.c_longlong, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "scasml_equation_g": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "scasml_equation_f": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "scasml_gp_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double, C.POINTER(C.c_void_p)]),
    "scasml_gp_destroy": (C.c_int, [C.c_void_p]),
    "scasml_gp_clone": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "scasml_gp_set_centres": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scasml_gp_set_alpha": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scasml_gp_get_alpha": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scasml_gp_gram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "scasml_gp_fit_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "scasml_gp_fit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                C.c_size_t, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p]),
    "scasml_gp_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scasml_gp_gradient_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_longlong]),
    "scasml_gp_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "scasml_picard_plan": (C.c_int, [C.POINTER(PicardParams), C.c_longlong, C.POINTER(C.c_size_t), C.POINTER(PicardStats)]),
    "scasml_uz_solve": (C.c_int, [C.c_void_p, C.POINTER(PicardParams), C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.POINTER(PicardStats), C.c_void_p]),
    "scasml_clip": (C.c_int, [C.c_void_p, C.c_longlong, C.c_double, C.c_void_p]),
    "scasml_gp_tc_supported": (C.c_int, [C.c_void_p]),
}

# exported by libscasml_b200_dbg.so only (tests/, tools/)
_DEBUG_SIGNATURES = {
    "scasml_debug_draw": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "scasml_debug_spd_inverse": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "scasml_debug_lu_solve": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "scasml_debug_tc_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]),
    "scasml_debug_tc_mma_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "scasml_debug_tc_pipe_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "scasml_debug_tc_timeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None
_lib_dbg = None
_table_devices = set()
_table_devices_dbg = set()


class ScasmlError(RuntimeError):
    pass


def load():
    """Load the compiled library; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScasmlError(
            f"{LIB_PATH} is missing: build it with `python -m scasml_gp_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_debug():
    """The debug build of the library (same sources + -DSCASML_DEBUG_HOOKS): test hooks and micro-benchmarks only.
    Handles created by the product library may be passed to it (identical structs); it has its own sampler table."""
    global _lib_dbg
    if _lib_dbg is not None:
        return _lib_dbg
    if not os.path.exists(LIB_DBG_PATH):
        raise ScasmlError(f"{LIB_DBG_PATH} is missing: build it with `python -m scasml_gp_b200.build`")
    lib = C.CDLL(LIB_DBG_PATH)
    for name, (res, args) in {**_SIGNATURES, **_DEBUG_SIGNATURES}.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib_dbg = lib
    return lib


def check(rc, lib=None):
    if rc == 0:
        return
    msg = (lib or load()).scasml_last_error()
    msg = msg.decode() if msg else "unknown error"
    if rc == ERR_NUMERIC:
        raise ValueError(msg)                      # mirrors models/GP.py:264-265
    raise ScasmlError(f"libscasml_b200 status {rc}: {msg}")


def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise ScasmlError("a CUDA device (B200, sm_100a) is required: this path has no CPU fallback")
    return torch


def stream_ptr():
    torch = torch_cuda()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def normal_half_table():
    """T[i] = float16(ndtri(0.5 + (i + 0.5)/65536)): the sampler's inverse-CDF table (see csrc/common.cuh)."""
    from scipy.special import ndtri
    i = np.arange(32768, dtype=np.float64)
    return ndtri(0.5 + (i + 0.5) / 65536.0).astype(np.float16)


def ensure_normal_table(debug=False):
    torch = torch_cuda()
    dev = torch.cuda.current_device()
    seen = _table_devices_dbg if debug else _table_devices
    if dev in seen:
        return
    tab = np.ascontiguousarray(normal_half_table().view(np.uint16))
    lib = load_debug() if debug else load()
    check(lib.scasml_set_normal_table(tab.ctypes.data_as(C.c_void_p)), lib)
    seen.add(dev)


_pin = {"buf": None, "evt": None}


def to_device(a, pinned=True):
    """NumPy (any float dtype) -> contiguous float64 CUDA tensor, staged through a cached pinned buffer (page-locking a fresh
    buffer per call costs more than the copy for the [B, d+1] inputs of a solve)."""
    torch = torch_cuda()
    if isinstance(a, torch.Tensor):                        # a pinned float64 host tensor goes to the device without a staging copy
        return a.to(device="cuda", dtype=torch.float64, non_blocking=bool(a.device.type == "cpu" and a.is_pinned())).contiguous()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if not (pinned and a.size) or a.size > (1 << 26):      # empty, or too large to page-lock a staging copy (> 512 MB)
        return torch.from_numpy(a).to("cuda")
    if _pin["evt"] is not None:
        _pin["evt"].synchronize()                      # the previous copy out of the staging buffer has finished
    if _pin["buf"] is None or _pin["buf"].numel() < a.size:
        _pin["buf"] = torch.empty(max(a.size, 1 << 16), dtype=torch.float64).pin_memory()
    stage = _pin["buf"][:a.size].view(a.shape)
    stage.copy_(torch.from_numpy(a))
    out = stage.to("cuda", non_blocking=True)
    _pin["evt"] = torch.cuda.Event()
    _pin["evt"].record()
    return out


_pin_out = {"buf": None}


def to_host(t):
    """CUDA tensor -> NumPy array through a cached pinned staging buffer (a pageable `.cpu()` goes through the driver's own
    bounce buffers and, with eight ranks on one host, serialises on them)."""
    torch = torch_cuda()
    t = t.contiguous()
    n = t.numel()
    if n == 0 or t.dtype != torch.float64 or n > (1 << 26):
        return t.cpu().numpy()
    if _pin_out["buf"] is None or _pin_out["buf"].numel() < n:
        _pin_out["buf"] = torch.empty(max(n, 1 << 16), dtype=torch.float64).pin_memory()
    stage = _pin_out["buf"][:n].view(t.shape)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return stage.numpy().copy()


def to_device_sharded(a, rank, world, dist):
    """Host -> device copy of an input every rank holds: each rank stages only its slice of the rows and the slices travel over NVLink
    (one all-gather) instead of `world` full copies through one host's staging buffers (sharded solves take the WHOLE batch on every rank)."""
    torch = torch_cuda()
    if isinstance(a, torch.Tensor) or world <= 1:
        return to_device(a)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    B = a.shape[0]
    per = -(-B // world)
    if a.ndim != 2 or B < 8 * world:
        return to_device(a)
    full = torch.empty((per * world, a.shape[1]), dtype=torch.float64, device="cuda")
    mine = a[rank * per:min(B, (rank + 1) * per)]
    part = full[rank * per:(rank + 1) * per]
    if mine.shape[0]:
        part[:mine.shape[0]].copy_(to_device(mine))
    if mine.shape[0] < per:
        part[mine.shape[0]:].zero_()
    dist.all_gather_into_tensor(full, part.clone())
    return full[:B]


class AbiComm(object):
    """The path's all-reduce through the C ABI (scasml_comm_*: NCCL loaded by the library) instead of torch.distributed -- what a host without
    PyTorch's process groups would do.  One process per GPU; rank 0 calls AbiComm.unique_id() and hands the 128 bytes to the other ranks.
    Assign to `solver.comm`; the solver then shards its sample units over `world` ranks and sums the partial blocks with all_reduce()."""

    class ReduceOp(object):
        SUM = "sum"

    def __init__(self, id_bytes, rank, world):
        self.rank, self.world = int(rank), int(world)
        torch_cuda()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(id_bytes))
        h = C.c_void_p()
        check(load().scasml_comm_init(buf, self.rank, self.world, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        check(load().scasml_comm_unique_id(buf))
        return bytes(buf)

    def all_reduce(self, t, op=None):
        assert t.is_cuda and t.is_contiguous() and str(t.dtype) == "torch.float64"
        check(load().scasml_allreduce_partial(self._h, ptr(t), t.numel(), stream_ptr()))

    def close(self):
        if getattr(self, "_h", None):
            load().scasml_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
