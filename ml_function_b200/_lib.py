"""ctypes binding of ``libkon_b200.so`` (the C-ABI declared in ``include/kon_b200.h``).

Tensors cross the boundary as DLPack ``DLTensor`` structs.  For torch tensors the
struct is filled directly from ``data_ptr()/shape/stride()`` (that *is* the DLPack
view of the tensor, without the capsule round trip); any other producer that speaks
``__dlpack__`` is accepted through its capsule.  There is no fallback of any kind:
if the shared library is missing or a call fails, a ``KonError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# KON_B200_LIB selects an experimental build of the same library (kernel-parameter sweeps)
LIB_PATH = os.environ.get("KON_B200_LIB") or os.path.join(_HERE, "libkon_b200.so")

KON_EMBED_SUM_FIELDS = 1
KON_EMBED_SKIP_INVALID = 2
KON_CIN_FP32, KON_CIN_BF16 = 0, 1
KON_ATTN_USE_SCALE, KON_ATTN_USE_LN, KON_ATTN_USE_RES, KON_ATTN_RELU, KON_ATTN_BF16 = 1, 2, 4, 8, 16


class KonError(RuntimeError):
    pass


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("device", DLDevice),
        ("ndim", ctypes.c_int32),
        ("dtype", DLDataType),
        ("shape", ctypes.POINTER(ctypes.c_int64)),
        ("strides", ctypes.POINTER(ctypes.c_int64)),
        ("byte_offset", ctypes.c_uint64),
    ]


class KonPut2D(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("src_pitch", ctypes.c_int64),
                ("dst_pitch", ctypes.c_int64), ("width", ctypes.c_int64), ("rows", ctypes.c_int64)]


class _DLManagedTensor(ctypes.Structure):
    _fields_ = [("dl_tensor", DLTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.c_void_p)]


_DTYPES = {
    torch.float32: (2, 32),
    torch.float64: (2, 64),
    torch.float16: (2, 16),
    torch.bfloat16: (4, 16),
    torch.int32: (0, 32),
    torch.int64: (0, 64),
    torch.uint8: (1, 8),
    torch.int8: (0, 8),
}

_DLT_P = ctypes.POINTER(DLTensor)


class TensorArg:
    """Keeps the shape/stride arrays alive for the duration of one call."""

    __slots__ = ("dl", "_shape", "_strides", "_keep")

    def __init__(self, t):
        if isinstance(t, torch.Tensor):
            nd = t.dim()
            self._shape = (ctypes.c_int64 * max(nd, 1))(*t.shape)
            self._strides = (ctypes.c_int64 * max(nd, 1))(*t.stride())
            code, bits = _DTYPES[t.dtype]
            dev = t.device
            self.dl = DLTensor(
                t.data_ptr(),
                DLDevice(2 if dev.type == "cuda" else 1, dev.index if dev.index is not None else 0),
                nd, DLDataType(code, bits, 1), self._shape, self._strides, 0)
            self._keep = t
        else:  # generic DLPack producer
            cap = t.__dlpack__()
            ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
            ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
            ptr = ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
            managed = ctypes.cast(ptr, ctypes.POINTER(_DLManagedTensor)).contents
            self.dl = managed.dl_tensor
            self._shape = self._strides = None
            self._keep = (t, cap)

    @property
    def ptr(self):
        return ctypes.byref(self.dl)


def _arg(t) -> Optional[TensorArg]:
    return None if t is None else TensorArg(t)


def _p(a: Optional[TensorArg]):
    return None if a is None else a.ptr


_lib = None


def lib():
    """Load libkon_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KonError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ml_function_b200/csrc`.  There is no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
    T = _DLT_P
    i64p = ctypes.POINTER(ctypes.c_int64)
    i32p = ctypes.POINTER(ctypes.c_int32)
    TT = ctypes.POINTER(_DLT_P)
    sigs = {
        "kon_abi_version": (ctypes.c_int, []),
        "kon_last_error": (ctypes.c_char_p, []),
        "kon_launch_count": (ctypes.c_longlong, []),
        "kon_profile_enable": (ctypes.c_int, [ctypes.c_int]),
        "kon_profile_reset": (ctypes.c_int, []),
        "kon_profile_read": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
        "kon_device_info": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)] + [ctypes.POINTER(ctypes.c_int)] * 2),
        "kon_embed_fwd": (ctypes.c_int, [T, T, i64p, i32, T, T, i32, vp]),
        "kon_embed_bwd_workspace_bytes": (sz, [i64, i32]),
        "kon_embed_bwd": (ctypes.c_int, [T, T, i64p, i32, T, T, T, T, vp]),
        "kon_embed_bwd_reuse": (ctypes.c_int, [T, T, i64p, i32, T, T, T, T, vp]),
        "kon_embed_bwd_pair": (ctypes.c_int, [T, T, T, i64p, i32, T, T, T, T, T, i32, vp]),
        "kon_embed_sort": (ctypes.c_int, [T, i64p, i32, T, vp]),
        "kon_embed_route_plan": (ctypes.c_int, [i64, i32, i32, i64p]),
        "kon_embed_sgd": (ctypes.c_int, [T, T, T, T, f32, f32, vp]),
        "kon_embed_adam": (ctypes.c_int, [T, T, T, T, T, T, f32, f32, f32, f32, f32, i32, vp]),
        "kon_embed_adam_devstep": (ctypes.c_int, [T, T, T, T, T, T, f32, f32, f32, f32, f32, T, vp]),
        "kon_peer_alloc": (ctypes.c_int, [ctypes.c_int, sz, ctypes.POINTER(vp), vp]),
        "kon_peer_open": (ctypes.c_int, [ctypes.c_int, vp, ctypes.POINTER(vp)]),
        "kon_peer_close": (ctypes.c_int, [ctypes.c_int, vp]),
        "kon_peer_free": (ctypes.c_int, [ctypes.c_int, vp]),
        "kon_peer_barrier": (ctypes.c_int, [ctypes.POINTER(vp), i32, i32, ctypes.c_int, i64, vp]),
        "kon_peer_put2d": (ctypes.c_int, [ctypes.POINTER(KonPut2D), i32, ctypes.c_int, vp]),
        "kon_embed_fwd_peer": (ctypes.c_int, [T, T, i64p, i32, ctypes.POINTER(vp), i32, i64, i64, i64, T, i32, vp]),
        "kon_embed_fwd_peer_cols": (ctypes.c_int, [T, T, i64p, i32, ctypes.POINTER(vp), i32, i64, i64, i64, i32p, T, i32, vp]),
        "kon_embed_bwd_peer": (ctypes.c_int, [ctypes.POINTER(vp), i32, i64, i64, i64, i32, T, i64p, i32, T, T, T, T, i32, vp]),
        "kon_fm_fwd": (ctypes.c_int, [T, T, T, vp]),
        "kon_fm_bwd": (ctypes.c_int, [T, T, T, T, vp]),
        "kon_fm_bwd_acc": (ctypes.c_int, [T, T, T, T, vp]),
        "kon_cross_fwd": (ctypes.c_int, [T, T, T, T, T, vp]),
        "kon_cross_bwd_workspace_bytes": (sz, [i64, i32, i32, ctypes.c_int]),
        "kon_cross_bwd": (ctypes.c_int, [T] * 9 + [vp]),
        "kon_cross_bwd_acc": (ctypes.c_int, [T] * 9 + [vp]),
        "kon_cin_saved_bytes": (sz, [i64, i32, i32, i32p, i32, i32]),
        "kon_cin_workspace_bytes": (sz, [i64, i32, i32, i32p, i32, i32, ctypes.c_int]),
        "kon_cin_fwd": (ctypes.c_int, [T, TT, TT, i32, T, T, T, i32, vp]),
        "kon_cin_bwd": (ctypes.c_int, [T, TT, TT, i32, T, T, T, TT, TT, T, i32, vp]),
        "kon_attn_fwd": (ctypes.c_int, [T] * 7 + [f32, i32, vp]),
        "kon_attn_bwd_workspace_bytes": (sz, [i64, i32, i32, i32, i32, ctypes.c_int]),
        "kon_attn_bwd": (ctypes.c_int, [T] * 14 + [f32, i32, vp]),
        "kon_head_fwd": (ctypes.c_int, [T] * 5 + [vp]),
        "kon_head_bwd_workspace_bytes": (sz, [i64, i32, i32, ctypes.c_int]),
        "kon_head_bwd": (ctypes.c_int, [T] * 9 + [vp]),
        "kon_pool_sum_fwd": (ctypes.c_int, [T, T, vp]),
        "kon_pairs_fwd": (ctypes.c_int, [T, T, vp]),
        "kon_pairs_bwd": (ctypes.c_int, [T, T, T, vp]),
        "kon_pattn_fwd": (ctypes.c_int, [T] * 5 + [i32, i32, vp]),
        "kon_pattn_bwd": (ctypes.c_int, [T] * 8 + [i32, i32, vp]),
        "kon_bce_workspace_bytes": (sz, []),
        "kon_bce_fwd": (ctypes.c_int, [T, T, T, T, f32, vp]),
        "kon_bce_bwd": (ctypes.c_int, [T, T, T, T, f32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)   # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if L.kon_abi_version() != 1:
        raise KonError(f"libkon_b200 ABI {L.kon_abi_version()} != 1")
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "kon_abi_version", "kon_last_error", "kon_launch_count", "kon_profile_enable", "kon_profile_reset",
    "kon_profile_read", "kon_device_info", "kon_embed_fwd",
    "kon_embed_bwd_workspace_bytes", "kon_embed_bwd", "kon_embed_bwd_reuse", "kon_embed_bwd_pair", "kon_embed_sort",
    "kon_embed_route_plan", "kon_embed_sgd", "kon_embed_adam",
    "kon_embed_adam_devstep",
    "kon_peer_alloc", "kon_peer_open", "kon_peer_close", "kon_peer_free", "kon_peer_barrier", "kon_peer_put2d",
    "kon_embed_fwd_peer", "kon_embed_fwd_peer_cols", "kon_embed_bwd_peer",
    "kon_fm_fwd", "kon_fm_bwd", "kon_fm_bwd_acc", "kon_cross_fwd", "kon_cross_bwd_workspace_bytes", "kon_cross_bwd",
    "kon_cross_bwd_acc",
    "kon_cin_saved_bytes", "kon_cin_workspace_bytes", "kon_cin_fwd", "kon_cin_bwd",
    "kon_attn_fwd", "kon_attn_bwd_workspace_bytes", "kon_attn_bwd",
    "kon_head_fwd", "kon_head_bwd_workspace_bytes", "kon_head_bwd",
    "kon_pool_sum_fwd", "kon_pairs_fwd", "kon_pairs_bwd", "kon_pattn_fwd", "kon_pattn_bwd",
    "kon_bce_workspace_bytes", "kon_bce_fwd", "kon_bce_bwd",
)


def profile_read(kernel: str):
    """-> (total ms, launches) of one kernel name recorded since kon_profile_reset()."""
    ms, n = ctypes.c_double(0.0), ctypes.c_longlong(0)
    check(lib().kon_profile_read(kernel.encode(), ctypes.byref(ms), ctypes.byref(n)), "kon_profile_read")
    return ms.value, n.value


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().kon_last_error()
        raise KonError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on ``device``.  For a non-CUDA device this is
    0 and the library's own validation rejects the tensors (KonError) -- no CPU path."""
    if device is not None and torch.device(device).type != "cuda":
        return 0
    return torch.cuda.current_stream(device).cuda_stream


def tensor_array(ts: Sequence[torch.Tensor]):
    """-> (ctypes array of DLTensor*, keepalive list)."""
    args = [TensorArg(t) for t in ts]
    arr = (_DLT_P * len(args))(*[ctypes.pointer(a.dl) for a in args])
    return arr, args


def i64_array(vals: Sequence[int]):
    return (ctypes.c_int64 * len(vals))(*vals)


def i32_array(vals: Sequence[int]):
    return (ctypes.c_int32 * len(vals))(*vals)
