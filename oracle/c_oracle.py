"""ctypes view of oracle/kon_oracle_c.c (plain-C restatement of the byte-exact parts of the hot path and of
the fp32 op order of FM / cross).  TEST INFRASTRUCTURE ONLY -- imported by tests/ (and compiled by
__graft_entry__.build()); never by the product.  Parity unpinned (see kon_oracle.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libkon_oracle_c.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "kon_oracle_c.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(_SO)
        f32p, i32p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_float, ctypes.c_int32, ctypes.c_int64))
        L.kon_c_embed_gather.argtypes = [f32p, i64p, i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, f32p]
        L.kon_c_embed_gather.restype = None
        L.kon_c_embed_grad.argtypes = [i64p, i32p, f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, i32p, f32p]
        L.kon_c_embed_grad.restype = ctypes.c_int64
        L.kon_c_fm.argtypes = [f32p, f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, f32p]
        L.kon_c_fm.restype = None
        L.kon_c_cross.argtypes = [f32p, f32p, f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, f32p]
        L.kon_c_cross.restype = None
        L.kon_c_cin.argtypes = [f32p, f32p, f32p, i32p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, f32p]
        L.kon_c_cin.restype = None
        L.kon_c_autoint_block.argtypes = [f32p] * 6 + [ctypes.c_int64] + [ctypes.c_int] * 4 + [f32p]
        L.kon_c_autoint_block.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def embed_gather(tables, offs, ids):
    tables, ids = _f32(tables), np.ascontiguousarray(ids, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    B, F = ids.shape
    k = tables.shape[1]
    out = np.empty((B, F, k), dtype=np.float32)
    lib().kon_c_embed_gather(_p(tables, ctypes.c_float), _p(offs, ctypes.c_int64), _p(ids, ctypes.c_int32), B, F, k,
                             _p(out, ctypes.c_float))
    return out


def embed_grad(offs, ids, d_out):
    ids, d_out = np.ascontiguousarray(ids, dtype=np.int32), _f32(d_out)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    B, F = ids.shape
    k = d_out.shape[2]
    rows = np.empty(B * F, dtype=np.int32)
    grads = np.empty((B * F, k), dtype=np.float32)
    n = lib().kon_c_embed_grad(_p(offs, ctypes.c_int64), _p(ids, ctypes.c_int32), _p(d_out, ctypes.c_float), B, F, k,
                               _p(rows, ctypes.c_int32), _p(grads, ctypes.c_float))
    return rows[:n], grads[:n]


def fm(v, lin=None):
    v = _f32(v)
    B, F, k = v.shape
    out = np.empty((B, k), dtype=np.float32)
    linp = None
    if lin is not None:
        lin = _f32(lin)
        linp = _p(lin, ctypes.c_float)
    lib().kon_c_fm(_p(v, ctypes.c_float), linp, B, F, k, _p(out, ctypes.c_float))
    return out


def cross(x0, w, b):
    x0, w, b = _f32(x0), _f32(w), _f32(b)
    B, D = x0.shape
    out = np.empty((B, D), dtype=np.float32)
    lib().kon_c_cross(_p(x0, ctypes.c_float), _p(w, ctypes.c_float), _p(b, ctypes.c_float), B, D, w.shape[0],
                      _p(out, ctypes.c_float))
    return out


def cin(x0, weights, biases):
    """x0 [B,m,D]; weights[l] [H_{l-1}*m, H_l]; biases[l] [H_l] -> pooled [B, L*D]."""
    x0 = _f32(x0)
    B, m, D = x0.shape
    hs = np.array([w.shape[1] for w in weights], dtype=np.int32)
    w = _f32(np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights]))
    b = _f32(np.concatenate([np.asarray(x, dtype=np.float32).reshape(-1) for x in biases]))
    out = np.empty((B, len(weights) * D), dtype=np.float32)
    lib().kon_c_cin(_p(x0, ctypes.c_float), _p(w, ctypes.c_float), _p(b, ctypes.c_float), _p(hs, ctypes.c_int32),
                    len(weights), B, m, D, _p(out, ctypes.c_float))
    return out


def autoint_block(x, wq, wk, wr, gamma, beta):
    """x [B,F,kin]; w* [kin,H,d] -> [H,B,F,d] (scaled scores, LayerNorm eps 1e-3, residual, ReLU)."""
    x, wq, wk, wr, gamma, beta = (_f32(a) for a in (x, wq, wk, wr, gamma, beta))
    B, F, kin = x.shape
    H, d = wq.shape[1], wq.shape[2]
    assert d <= 64
    y = np.empty((H, B, F, d), dtype=np.float32)
    lib().kon_c_autoint_block(*[_p(a, ctypes.c_float) for a in (x, wq, wk, wr, gamma, beta)], B, F, kin, H, d,
                              _p(y, ctypes.c_float))
    return y
