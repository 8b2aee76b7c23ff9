"""torch.autograd bindings of the C-ABI kernels (one Function per reference ``call``).

Forward = the ``kon_*_fwd`` entry point, backward = ``kon_*_bwd``; torch only owns
the buffers and the stream.  Nothing here computes on the CPU or with torch ops:
inputs that are not CUDA tensors raise ``KonError`` from the library's own checks.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L


# Per-op device timing for bench.py's roofline: when PROFILE is a dict, every C-ABI call below
# is bracketed by CUDA events recorded on the launching (current) stream.
PROFILE = None


# KON_NVTX=1: every C-ABI call is wrapped in an NVTX range named after the op (nsys / ncu --nvtx timelines).
NVTX = os.environ.get("KON_NVTX", "0") == "1"


class _prof:
    __slots__ = ("name", "e0")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push("kon." + self.name)
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.setdefault(self.name, []).append((self.e0, e1))
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def profile_summary():
    """-> {op: (calls, mean ms)} from the events collected in PROFILE (synchronises)."""
    torch.cuda.synchronize()
    out = {}
    for k, evs in (PROFILE or {}).items():
        ts = [a.elapsed_time(b) for a, b in evs]
        out[k] = (len(ts), sum(ts) / max(len(ts), 1))
    return out


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- #
# embeddings (a1-a4)
# --------------------------------------------------------------------------- #
class SparseGrad:
    """Result of kon_embed_bwd: ``rows[:n]`` ascending unique arena rows and their summed
    gradients ``grads[:n]``; ``n`` stays on the device (no host sync)."""

    __slots__ = ("rows", "grads", "n", "disjoint")

    def __init__(self, rows, grads, n, disjoint=False):
        self.rows, self.grads, self.n = rows, grads, n
        self.disjoint = disjoint        # True: shares no row with the other gradients of its arena in this step

    def to_dense(self, n_rows: int) -> torch.Tensor:
        n = int(self.n.item())
        dense = torch.zeros(n_rows, self.grads.shape[1], dtype=self.grads.dtype, device=self.grads.device)
        dense.index_add_(0, self.rows[:n].long(), self.grads[:n])
        return dense


def embed_fwd_raw(arena, ids, field_row_offset: Sequence[int], sum_fields=False, out=None, oob=None):
    lib = L.lib()
    F = ids.shape[1]
    dim = arena.shape[1]
    if out is None:
        shape = (ids.shape[0], dim) if sum_fields else (ids.shape[0], F, dim)
        out = torch.empty(shape, dtype=arena.dtype, device=arena.device)
    offs = L.i64_array(list(field_row_offset))
    a, i, o, ob = L._arg(arena), L._arg(ids), L._arg(out), L._arg(oob)
    with _prof("embed_fwd" if dim > 1 else "embed_fwd_lin"):
        L.check(lib.kon_embed_fwd(a.ptr, i.ptr, offs, F, o.ptr, L._p(ob),
                                  L.KON_EMBED_SUM_FIELDS if sum_fields else 0, L.stream_ptr(arena.device)),
                "kon_embed_fwd")
    return out


# Routing (sort) results of the current step, keyed by (ids storage, offsets): the embedding and the
# first-order tables of a model are looked up with the same ids and the same per-field row counts, so
# the second backward reuses the first one's sorted keys (kon_embed_bwd_reuse).  Cleared by new_step().
_SORT_CACHE = {}
_SHARE_SORT = False      # only inside new_step() ... end_step(): the ids tensors are alive (saved for the
                         # backward), so a storage address identifies them


_STEP_CACHE = {}         # other per-step results keyed the same way (parallel.py: the exchanged ids)


def _join_side_streams():
    # a routing sort nobody consumed (forward-only step): rejoin, so that no side-stream work outlives the step
    for key, ev in list(_SORT_EVENTS.items()):
        torch.cuda.current_stream().wait_event(ev)
    _SORT_EVENTS.clear()


_STEP_PRESORT = True


# Gradient of the concat buffer, accumulated in place.  ``xcat [B,W]`` feeds several branches (MLP + FM in DeepFM,
# MLP + cross in DCN); autograd would sum their full-width gradients with an extra pass over two [B,W] tensors
# (43 us for DeepFM, ~100 us for DCN at B = 65,536).  Inside new_step() ... end_step() the first branch whose backward
# produces a full-width fp32 gradient (the first Dense layer: its dgrad GEMM output) offers that buffer here; a later
# branch finds it, adds its own gradient INTO it (kon_fm_bwd_acc / kon_cross_bwd_acc) and returns None to autograd.
# Whichever order autograd picks is fine: a branch that finds nothing returns its gradient the ordinary way.
# Only buffers a builder has declared to have EXACTLY two consumers take part (``two_branch_input``): with a third
# consumer autograd would already have summed the offered gradient into a new tensor when the late branch adds into
# the old one, and that contribution would be lost.
ACC_XGRAD = os.environ.get("KON_ACC_XGRAD", "1") != "0"
_XGRAD = {}
_XGRAD_OK = set()


def two_branch_input(x: torch.Tensor) -> torch.Tensor:
    """Declare that ``x`` (the concat buffer) feeds exactly two branches in this forward -- the MLP, whose first Dense
    layer offers its input gradient, and ONE of FM / cross, which adds into it (DeepFM MD:80-90, DCN MD:92-106)."""
    if ACC_XGRAD and _SHARE_SORT:
        _XGRAD_OK.add(xgrad_key(x))
    return x


def offer_xgrad(key, g: torch.Tensor):
    if ACC_XGRAD and _SHARE_SORT and key in _XGRAD_OK and g is not None and g.dtype == torch.float32 \
            and g.is_contiguous():
        _XGRAD[key] = g


def take_xgrad(key, shape):
    g = _XGRAD.pop(key, None) if (ACC_XGRAD and _SHARE_SORT) else None      # one taker per offer
    if g is not None and tuple(g.shape) == tuple(shape):
        return g
    return None


def xgrad_key(x: torch.Tensor):
    return (x.data_ptr(), tuple(x.shape))


def new_step(presort: bool = True):
    """``presort=False``: keep the routing sort in the backward (steps dominated by the persistent tcgen05 CIN
    kernels: sort kernels co-scheduled with them cost more than they hide, 9.95 -> 10.11 ms measured)."""
    global _SHARE_SORT, _STEP_PRESORT
    _join_side_streams()
    _SORT_CACHE.clear()
    _STEP_CACHE.clear()
    _XGRAD.clear()
    _XGRAD_OK.clear()
    _SHARE_SORT = True
    _STEP_PRESORT = presort


def end_step():
    global _SHARE_SORT
    # sharded jobs: a first-order gradient whose owner-side scatter was deferred to the embedding backward's barrier
    # (parallel._ShardedSumPeer) and never picked up -- publish and scatter it now (same on every rank)
    for key in [k_ for k_ in _STEP_CACHE if isinstance(k_, tuple) and k_ and k_[0] == "lin_bwd"]:
        _STEP_CACHE.pop(key)(need_barrier=True)
    flush_deferred()
    _join_side_streams()
    _SORT_CACHE.clear()
    _STEP_CACHE.clear()
    _XGRAD.clear()
    _XGRAD_OK.clear()
    _SHARE_SORT = False


def _bwd_workspace(ids, field_row_offset, n, dim, dev, share_sort):
    """-> (workspace, reuse): ``reuse`` when an earlier backward of this step left the sorted routing of the
    same ids / offsets at the front of the returned workspace."""
    lib = L.lib()
    if share_sort is None:
        share_sort = _SHARE_SORT
    key = (ids.data_ptr(), ids._version, tuple(field_row_offset), n)
    cached = _SORT_CACHE.get(key) if share_sort else None
    need = lib.kon_embed_bwd_workspace_bytes(n, dim)
    if cached is not None:
        ev = _SORT_EVENTS.pop(key, None)
        if ev is not None:               # routed on the side stream: join it once -- also before the buffer may be
            torch.cuda.current_stream(dev).wait_event(ev)      # dropped below (the sort may still be writing it)
        if cached.numel() >= need:
            return cached, True
    # sized for the widest payload seen in practice plus the dim-1 path, so a later call can reuse it
    ws = _ws(max(need, lib.kon_embed_bwd_workspace_bytes(n, 1), lib.kon_embed_bwd_workspace_bytes(n, 32)), dev)
    if share_sort:
        _SORT_CACHE[key] = ws
    return ws, False


# The routing of the backward (per-field counting sort, run-head count: ~0.1 ms of latency-bound launches for 1.7 M
# lookups) depends on the ids alone.  Inside a training step it is therefore issued at the START of the step on
# a side stream (kon_embed_sort), overlapped with the forward / interaction kernels; the backward waits on its
# event and runs only the segmented reduction (kon_embed_bwd_reuse).  KON_PRESORT=0 keeps it in the backward.
_SIDE_STREAMS = {}
_SORT_EVENTS = {}
PRESORT = os.environ.get("KON_PRESORT", "1") != "0"


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def join_presort():
    """Make the current stream wait for the routing sorts in flight on the side stream.  Called before a kernel
    that owns the whole GPU (the persistent tcgen05 CIN kernels: one CTA per SM, ~200 KB of shared memory): a sort
    still running when such a kernel launches keeps some SMs busy, those CTAs start late and the whole persistent
    kernel ends late -- worse than not overlapping at all."""
    for key, ev in list(_SORT_EVENTS.items()):
        torch.cuda.current_stream().wait_event(ev)
    _SORT_EVENTS.clear()


def embed_presort(ids, field_row_offset: Sequence[int]):
    """Start the routing sort for ``ids`` on the side stream (no-op outside new_step() ... end_step(), or when
    the same ids / offsets were already routed in this step)."""
    if not (_SHARE_SORT and PRESORT and _STEP_PRESORT) or not ids.is_cuda or ids.numel() == 0:
        return
    lib = L.lib()
    n = ids.numel()
    key = (ids.data_ptr(), ids._version, tuple(field_row_offset), n)
    if key in _SORT_CACHE:
        return
    dev = ids.device
    ws = _ws(max(lib.kon_embed_bwd_workspace_bytes(n, 1), lib.kon_embed_bwd_workspace_bytes(n, 32)), dev)
    cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
    side.wait_stream(cur)                    # the ids (H2D copy, exchange) are produced on the current stream
    offs = L.i64_array(list(field_row_offset))
    a, w = L._arg(ids), L._arg(ws)
    with torch.cuda.stream(side):
        L.check(lib.kon_embed_sort(a.ptr, offs, ids.shape[1], w.ptr, side.cuda_stream), "kon_embed_sort")
        ev = torch.cuda.Event()
        ev.record(side)
    # no record_stream needed: `ws` (held by _SORT_CACHE) and `ids` (saved for the backward) stay alive until
    # the side stream has been joined back into the current stream (wait_event in the backward / end_step)
    _SORT_CACHE[key] = ws
    _SORT_EVENTS[key] = ev


def embed_bwd_raw(d_out, ids, field_row_offset: Sequence[int], share_sort: Optional[bool] = None, lin=None):
    """d_out [B,F,dim] (any strides on dims 0/1, e.g. an expanded [B,1,dim]) -> SparseGrad.
    ``lin``: a second gradient [B,F,1] (any strides) of dim-1 tables looked up with the same ids / offsets, reduced
    in the same pass (kon_embed_bwd_pair) -> (SparseGrad, SparseGrad of the dim-1 tables; rows and n shared)."""
    lib = L.lib()
    F = ids.shape[1]
    dim = d_out.shape[2]
    n = ids.numel()
    dev = d_out.device
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    grads = torch.empty((max(n, 1), dim), dtype=torch.float32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    ws, reuse = _bwd_workspace(ids, field_row_offset, n, dim, dev, share_sort)
    if lin is not None:
        grads1 = torch.empty((max(n, 1), 1), dtype=torch.float32, device=dev)
        offs = L.i64_array(list(field_row_offset))
        a = [L._arg(t) for t in (d_out, lin, ids, rows, grads, grads1, nu, ws)]
        with _prof("embed_bwd"):
            L.check(lib.kon_embed_bwd_pair(a[0].ptr, a[1].ptr, a[2].ptr, offs, F, a[3].ptr, a[4].ptr, a[5].ptr, a[6].ptr,
                                           a[7].ptr, 1 if reuse else 0, L.stream_ptr(dev)), "kon_embed_bwd_pair")
        return SparseGrad(rows, grads, nu), SparseGrad(rows, grads1, nu)
    fn, what = (lib.kon_embed_bwd_reuse, "kon_embed_bwd_reuse") if reuse else (lib.kon_embed_bwd, "kon_embed_bwd")
    offs = L.i64_array(list(field_row_offset))
    a = [L._arg(t) for t in (d_out, ids, rows, grads, nu, ws)]
    with _prof("embed_bwd" if dim > 1 else "embed_bwd_lin"):
        L.check(fn(a[0].ptr, a[1].ptr, offs, F, a[2].ptr, a[3].ptr, a[4].ptr, a[5].ptr,
                   L.stream_ptr(dev)), what)
    return SparseGrad(rows, grads, nu)


# ---- sharded embeddings over NVLink peer memory ----------------------------------
def embed_fwd_peer(arena, ids, field_row_offset: Sequence[int], peer_out, n_peers: int, rows_per_peer: int,
                   stride_b: int, stride_f: int, skip_invalid=False, oob=None, field_col: Optional[Sequence[int]] = None):
    """kon_embed_fwd_peer: gather this rank's tables for the GLOBAL batch ``ids`` [B_g, F_loc] and store
    every row into the buffer of the rank that owns the sample.  ``peer_out``: ctypes ``c_void_p`` array
    (this process's mappings of the ranks' buffers, already offset to this rank's first column)."""
    lib = L.lib()
    offs = L.i64_array(list(field_row_offset))
    a, i, ob = L._arg(arena), L._arg(ids), L._arg(oob)
    cols = None if field_col is None else L.i32_array(list(field_col))      # float offset of each local field's column
    with _prof("embed_fwd_peer"):
        L.check(lib.kon_embed_fwd_peer_cols(a.ptr, i.ptr, offs, ids.shape[1], peer_out, n_peers, rows_per_peer,
                                            stride_b, stride_f, cols, L._p(ob),
                                            L.KON_EMBED_SKIP_INVALID if skip_invalid else 0,
                                            L.stream_ptr(arena.device)), "kon_embed_fwd_peer_cols")


def embed_bwd_peer(peer_d_out, n_peers: int, rows_per_peer: int, stride_b: int, stride_f: int, dim: int,
                   ids, field_row_offset: Sequence[int], share_sort: Optional[bool] = None) -> SparseGrad:
    """kon_embed_bwd_peer: sort-then-segment scatter-add whose gradient rows are loaded from the ranks
    that produced them (``peer_d_out``: ctypes ``c_void_p`` array of mapped gradient buffers)."""
    lib = L.lib()
    n = ids.numel()
    dev = ids.device
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    grads = torch.empty((max(n, 1), dim), dtype=torch.float32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    ws, reuse = _bwd_workspace(ids, field_row_offset, n, dim, dev, share_sort)
    offs = L.i64_array(list(field_row_offset))
    a = [L._arg(t) for t in (ids, rows, grads, nu, ws)]
    with _prof("embed_bwd_peer"):
        L.check(lib.kon_embed_bwd_peer(peer_d_out, n_peers, rows_per_peer, stride_b, stride_f, dim, a[0].ptr, offs,
                                       ids.shape[1], a[1].ptr, a[2].ptr, a[3].ptr, a[4].ptr, 1 if reuse else 0,
                                       L.stream_ptr(dev)), "kon_embed_bwd_peer")
    return SparseGrad(rows, grads, nu)


# The first-order ("linear", dim-1) tables of a model are looked up with the same ids as its embedding tables
# (FeatureInput(useLinear=True), DP:65-76).  Inside new_step() ... end_step() their gradient -- which autograd
# delivers FIRST, the linear term sits at the end of the forward -- is parked until the embedding tables' backward
# and reduced in the same pass over the shared routing (kon_embed_bwd_pair).  ``flush_deferred()`` (Trainer: right
# after backward(); also end_step()) scatters whatever was parked and never picked up.
FUSE_LIN = os.environ.get("KON_FUSE_LIN", "1") != "0"


def _route_key(ids, field_row_offset):
    return (ids.data_ptr(), ids._version, tuple(field_row_offset), ids.numel())


def _announce_main(arena, ids, field_row_offset):
    if FUSE_LIN and _SHARE_SORT and arena.requires_grad and ids.dim() == 2 and arena.shape[1] % 4 == 0:
        _STEP_CACHE[("emb_main", _route_key(ids, field_row_offset))] = True


def _append_grad(arena, sg):
    if not hasattr(arena, "kon_sparse_grads"):
        arena.kon_sparse_grads = []
    arena.kon_sparse_grads.append(sg)


def _main_backward(arena, g3, ids, field_row_offset):
    """The embedding tables' scatter-add, taking a parked first-order gradient of the same routing along."""
    key = _route_key(ids, field_row_offset)
    _STEP_CACHE.pop(("emb_main", key), None)
    parked = _STEP_CACHE.pop(("lin_parked", key), None)
    if parked is None:
        _append_grad(arena, embed_bwd_raw(g3, ids, field_row_offset))
        return
    lin_arena, g1 = parked
    sg, sg1 = embed_bwd_raw(g3, ids, field_row_offset, lin=g1)
    _append_grad(arena, sg)
    _append_grad(lin_arena, sg1)


def flush_deferred():
    for key in [k_ for k_ in _STEP_CACHE if isinstance(k_, tuple) and k_ and k_[0] == "lin_parked"]:
        lin_arena, g1 = _STEP_CACHE.pop(key)
        ids = _STEP_CACHE.pop(("lin_ids", key[1]))
        _append_grad(lin_arena, embed_bwd_raw(g1, ids, key[1][2]))


class _EmbedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, arena, ids, field_row_offset, sum_fields):
        if arena.requires_grad:
            embed_presort(ids, field_row_offset)
            if arena.shape[1] > 1:
                _announce_main(arena, ids, field_row_offset)
        out = embed_fwd_raw(arena.detach(), ids, field_row_offset, sum_fields)
        ctx.save_for_backward(ids)
        ctx.arena = arena
        ctx.offs = field_row_offset
        ctx.sum_fields = sum_fields
        return out

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        arena = ctx.arena
        if arena.requires_grad:
            if ctx.sum_fields:   # [B,dim] -> every field sees the same gradient (stride 0)
                g3 = g.contiguous().unsqueeze(1).expand(g.shape[0], ids.shape[1], g.shape[1])
            else:
                g3 = g.contiguous()
            key = _route_key(ids, ctx.offs)
            if arena.shape[1] == 1 and ids.dim() == 2 and ("emb_main", key) in _STEP_CACHE \
                    and ("lin_parked", key) not in _STEP_CACHE:
                _STEP_CACHE[("lin_parked", key)] = (arena, g3)       # reduced by the embedding tables' backward
                _STEP_CACHE[("lin_ids", key)] = ids
            elif arena.shape[1] > 1 and g3.dim() == 3 and g3.stride(2) == 1 and g3.stride(0) % 4 == 0 \
                    and g3.stride(1) % 4 == 0 and g3.data_ptr() % 16 == 0:
                _main_backward(arena, g3, ids, ctx.offs)
            else:
                _append_grad(arena, embed_bwd_raw(g3, ids, ctx.offs))
        return None, None, None, None


def embed_lookup(arena, ids, field_row_offset, sum_fields=False):
    """Gather (and bag-sum) rows of ``arena``; the gradient of ``arena`` is delivered
    sparsely as ``arena.kon_sparse_grads`` (list of SparseGrad) and ``arena.grad`` stays
    None -- a dense [R,dim] gradient is never materialised."""
    return _EmbedLookup.apply(arena, ids, tuple(field_row_offset), sum_fields)


def embed_sgd(arena, sg: SparseGrad, lr: float, l2: float = 0.0):
    lib = L.lib()
    a = [L._arg(t) for t in (arena, sg.rows, sg.grads, sg.n)]
    with _prof("embed_sgd"):
        L.check(lib.kon_embed_sgd(a[0].ptr, a[1].ptr, a[2].ptr, a[3].ptr, lr, l2, L.stream_ptr(arena.device)),
                "kon_embed_sgd")


def embed_adam_devstep(arena, m, v, sg: SparseGrad, lr, beta1, beta2, eps, l2, step_dev):
    """Row-wise lazy Adam with the step counter on the device (CUDA-graph friendly)."""
    lib = L.lib()
    a = [L._arg(t) for t in (arena, m, v, sg.rows, sg.grads, sg.n, step_dev)]
    with _prof("embed_adam"):
        L.check(lib.kon_embed_adam_devstep(a[0].ptr, a[1].ptr, a[2].ptr, a[3].ptr, a[4].ptr, a[5].ptr, lr, beta1,
                                           beta2, eps, l2, a[6].ptr, L.stream_ptr(arena.device)),
                "kon_embed_adam_devstep")


def embed_adam(arena, m, v, sg: SparseGrad, lr, beta1, beta2, eps, l2, step):
    lib = L.lib()
    a = [L._arg(t) for t in (arena, m, v, sg.rows, sg.grads, sg.n)]
    with _prof("embed_adam"):
        L.check(lib.kon_embed_adam(a[0].ptr, a[1].ptr, a[2].ptr, a[3].ptr, a[4].ptr, a[5].ptr, lr, beta1,
                                   beta2, eps, l2, step, L.stream_ptr(arena.device)), "kon_embed_adam")


# --------------------------------------------------------------------------- #
# FM (a5-a6)
# --------------------------------------------------------------------------- #
class _Fm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, lin):
        lib = L.lib()
        out = torch.empty((v.shape[0], v.shape[2]), dtype=v.dtype, device=v.device)
        a, b, o = L._arg(v), L._arg(lin), L._arg(out)
        with _prof("fm_fwd"):
            L.check(lib.kon_fm_fwd(a.ptr, L._p(b), o.ptr, L.stream_ptr(v.device)), "kon_fm_fwd")
        ctx.save_for_backward(v)
        ctx.has_lin = lin is not None
        ctx.lin_shape = None if lin is None else lin.shape
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        (v,) = ctx.saved_tensors
        g = g.contiguous()
        dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
        dlin = torch.empty(ctx.lin_shape, dtype=v.dtype, device=v.device) if ctx.has_lin else None
        a, b, c, d = L._arg(v), L._arg(g), L._arg(dv), L._arg(dlin)
        with _prof("fm_bwd"):
            L.check(lib.kon_fm_bwd(a.ptr, b.ptr, c.ptr, L._p(d), L.stream_ptr(v.device)), "kon_fm_bwd")
        return dv, dlin


class _FmXcat(torch.autograd.Function):
    """FM on the fields window of the concat buffer ``xcat [B,W]`` (columns ``[0,F*k)``): the backward
    writes ``dv`` straight into a ``[B,W]`` gradient of ``xcat`` (strided output of kon_fm_bwd) instead of
    autograd's slice_backward (a zero fill plus a copy of the same 109 MB)."""

    @staticmethod
    def forward(ctx, xcat, lin, F, k):
        lib = L.lib()
        v = xcat[:, :F * k].view(xcat.shape[0], F, k)
        out = torch.empty((xcat.shape[0], k), dtype=xcat.dtype, device=xcat.device)
        a, b, o = L._arg(v), L._arg(lin), L._arg(out)
        with _prof("fm_fwd"):
            L.check(lib.kon_fm_fwd(a.ptr, L._p(b), o.ptr, L.stream_ptr(xcat.device)), "kon_fm_fwd")
        ctx.save_for_backward(xcat)
        ctx.F, ctx.k = F, k
        ctx.lin_shape = None if lin is None else lin.shape
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        (xcat,) = ctx.saved_tensors
        F, k = ctx.F, ctx.k
        B, W = xcat.shape
        g = g.contiguous()
        v = xcat[:, :F * k].view(B, F, k)
        dlin = torch.empty(ctx.lin_shape, dtype=xcat.dtype, device=xcat.device) if ctx.lin_shape is not None else None
        acc = take_xgrad(xgrad_key(xcat), (B, W))
        if acc is not None:           # another branch's gradient of xcat is already there: add into it
            dv = acc[:, :F * k].view(B, F, k)
            a, b, c, d = L._arg(v), L._arg(g), L._arg(dv), L._arg(dlin)
            with _prof("fm_bwd"):
                L.check(lib.kon_fm_bwd_acc(a.ptr, b.ptr, c.ptr, L._p(d), L.stream_ptr(xcat.device)), "kon_fm_bwd_acc")
            return None, dlin, None, None
        gx = torch.empty((B, W), dtype=xcat.dtype, device=xcat.device)
        if W > F * k:
            gx[:, F * k:].zero_()
        dv = gx[:, :F * k].view(B, F, k)
        a, b, c, d = L._arg(v), L._arg(g), L._arg(dv), L._arg(dlin)
        with _prof("fm_bwd"):
            L.check(lib.kon_fm_bwd(a.ptr, b.ptr, c.ptr, L._p(d), L.stream_ptr(xcat.device)), "kon_fm_bwd")
        return gx, dlin, None, None


def fm_xcat(xcat: torch.Tensor, lin: Optional[torch.Tensor], F: int, k: int) -> torch.Tensor:
    """FM over ``xcat[:, :F*k]`` viewed as ``[B,F,k]``; gradient delivered as a full ``[B,W]`` tensor."""
    return _FmXcat.apply(xcat, lin, F, k)


def fm(v: torch.Tensor, lin: Optional[torch.Tensor] = None) -> torch.Tensor:
    """v [B,F,k], lin [B,F] or None -> [B,k]  (FmLayer / InnerLayer(use_add=True))."""
    return _Fm.apply(v, lin)


# --------------------------------------------------------------------------- #
# DCN cross (a7)
# --------------------------------------------------------------------------- #
class _Cross(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, w, b):
        lib = L.lib()
        w, b = w.contiguous(), b.contiguous()
        if x0.stride(-1) != 1:
            x0 = x0.contiguous()
        out = torch.empty(x0.shape, dtype=x0.dtype, device=x0.device)
        s = torch.empty((x0.shape[0], w.shape[0]), dtype=x0.dtype, device=x0.device)
        a = [L._arg(t) for t in (x0, w, b, out, s)]
        with _prof("cross_fwd"):
            L.check(lib.kon_cross_fwd(*[t.ptr for t in a], L.stream_ptr(x0.device)), "kon_cross_fwd")
        ctx.save_for_backward(x0, w, b, s)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        x0, w, b, s = ctx.saved_tensors
        if g.stride(-1) != 1:
            g = g.contiguous()
        dev = x0.device
        acc = take_xgrad(xgrad_key(x0), x0.shape)        # another branch's gradient of x0 (the concat buffer)
        dx0 = acc if acc is not None else torch.empty(x0.shape, dtype=x0.dtype, device=dev)
        dw = torch.empty_like(w)
        db = torch.empty_like(b)
        ws = _ws(lib.kon_cross_bwd_workspace_bytes(x0.shape[0], x0.shape[1], w.shape[0], dev.index or 0), dev)
        a = [L._arg(t) for t in (x0, w, b, s, g, dx0, dw, db, ws)]
        fn, what = (lib.kon_cross_bwd_acc, "kon_cross_bwd_acc") if acc is not None else (lib.kon_cross_bwd, "kon_cross_bwd")
        with _prof("cross_bwd"):
            L.check(fn(*[t.ptr for t in a], L.stream_ptr(dev)), what)
        return (None if acc is not None else dx0), dw, db


def cross(x0, w, b):
    """x0 [B,D], w/b [L,D] -> x_L [B,D]  (CrossLayer)."""
    return _Cross.apply(x0, w, b)


# --------------------------------------------------------------------------- #
# CIN (a8)
# --------------------------------------------------------------------------- #
class _Cin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, precision, n_layers, fields, *wb):
        """x0: [B,m,D], or (fields=(m,D)) the concat buffer xcat [B,W] whose first m*D columns are
        the field embeddings -- the bf16 path then reads them in place (batch stride W) and the
        backward returns the gradient of the whole buffer, written in place too."""
        lib = L.lib()
        ws_ = [t.contiguous() for t in wb[:n_layers]]
        bs_ = [t.contiguous() for t in wb[n_layers:]]
        src = x0
        if fields is not None:
            m, D = fields
            x0 = x0[:, :m * D].view(x0.shape[0], m, D)
        strided_ok = precision == L.KON_CIN_BF16 and x0.stride(2) == 1 and x0.stride(1) == x0.shape[2] \
            and x0.stride(0) % 4 == 0 and x0.data_ptr() % 16 == 0
        if not strided_ok:
            x0 = x0.contiguous()
        B, m, D = x0.shape
        dev = x0.device
        hs = L.i32_array([w.shape[1] for w in ws_])
        pooled = torch.empty((B, n_layers * D), dtype=torch.float32, device=dev)
        saved = _ws(lib.kon_cin_saved_bytes(B, m, D, hs, n_layers, precision), dev)
        work = _ws(lib.kon_cin_workspace_bytes(B, m, D, hs, n_layers, precision, dev.index or 0), dev)
        wa, wk = L.tensor_array(ws_)
        ba, bk = L.tensor_array(bs_)
        a = [L._arg(t) for t in (x0, pooled, saved, work)]
        with _prof("cin_fwd"):
            L.check(lib.kon_cin_fwd(a[0].ptr, wa, ba, n_layers, a[1].ptr, a[2].ptr, a[3].ptr, precision,
                                    L.stream_ptr(dev)), "kon_cin_fwd")
        ctx.save_for_backward(x0, saved, *ws_, *bs_)
        ctx.precision, ctx.n_layers = precision, n_layers
        ctx.work = work
        ctx.fields = fields
        ctx.src_shape = tuple(src.shape)
        return pooled

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        x0, saved = ctx.saved_tensors[:2]
        nl = ctx.n_layers
        ws_ = list(ctx.saved_tensors[2:2 + nl])
        bs_ = list(ctx.saved_tensors[2 + nl:])
        dev = x0.device
        g = g.contiguous()
        B, m, D = x0.shape
        if ctx.fields is not None and ctx.precision == L.KON_CIN_BF16:
            gsrc = torch.empty(ctx.src_shape, dtype=x0.dtype, device=dev)     # gradient of xcat
            if ctx.src_shape[1] > m * D:
                gsrc[:, m * D:] = 0
            dx0 = gsrc[:, :m * D].view(B, m, D)
        else:
            dx0 = torch.empty((B, m, D), dtype=x0.dtype, device=dev)
            gsrc = dx0
            if ctx.fields is not None:
                gsrc = torch.zeros(ctx.src_shape, dtype=x0.dtype, device=dev)
        dws = [torch.empty_like(w) for w in ws_]
        dbs = [torch.empty_like(b) for b in bs_]
        wa, k1 = L.tensor_array(ws_)
        ba, k2 = L.tensor_array(bs_)
        dwa, k3 = L.tensor_array(dws)
        dba, k4 = L.tensor_array(dbs)
        a = [L._arg(t) for t in (x0, g, saved, dx0, ctx.work)]
        with _prof("cin_bwd"):
            L.check(lib.kon_cin_bwd(a[0].ptr, wa, ba, nl, a[1].ptr, a[2].ptr, a[3].ptr, dwa, dba, a[4].ptr,
                                    ctx.precision, L.stream_ptr(dev)), "kon_cin_bwd")
        if ctx.fields is not None and gsrc is not dx0 and gsrc.data_ptr() != dx0.data_ptr():
            gsrc[:, :m * D] = dx0.reshape(B, m * D)
        return (gsrc, None, None, None, *dws, *dbs)


def cin(x0, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], precision: int = L.KON_CIN_FP32,
        fields=None):
    """x0 [B,m,D] (or, with ``fields=(m,D)``, the concat buffer [B,W] holding the fields in its first
    m*D columns), weights[l] [H_{l-1}*m, H_l], biases[l] [H_l] -> pooled [B, L*D]."""
    return _Cin.apply(x0, precision, len(weights), fields, *weights, *biases)


# --------------------------------------------------------------------------- #
# AutoInt attention (a9-a10)
# --------------------------------------------------------------------------- #
# Memory layouts of the attention output.  The result is always the [H,B,F,d] tensor of BL:377; with
# "bfhd" / "bhfd" it is a permuted window of a buffer laid out [B,F,H,d] / [B,H,F,d], so that the consumer's
# re-packing (the next attention layer reads [B,F,H*d]; MergeScoreLayer reads the heads side by side as
# [B,H*F*d], MD:162) is a free view instead of a permute copy, forward and backward (bf16 path only).
_ATTN_LAYOUTS = {"hbfd": None, "bfhd": (2, 0, 1, 3), "bhfd": (1, 0, 2, 3)}


class _Attn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wq, wk, wr, gamma, beta, flags, eps, layout):
        lib = L.lib()
        x, wq, wk = x.contiguous(), wq.contiguous(), wk.contiguous()
        wr = None if wr is None else wr.contiguous()
        H, d = wq.shape[1], wq.shape[2]
        B, F = x.shape[0], x.shape[1]
        if layout == "bfhd":
            y = torch.empty((B, F, H, d), dtype=x.dtype, device=x.device).permute(2, 0, 1, 3)
        elif layout == "bhfd":
            y = torch.empty((B, H, F, d), dtype=x.dtype, device=x.device).permute(1, 0, 2, 3)
        else:
            y = torch.empty((H, B, F, d), dtype=x.dtype, device=x.device)
        a = [L._arg(t) for t in (x, wq, wk, wr, gamma, beta, y)]
        with _prof("attn_fwd"):
            L.check(lib.kon_attn_fwd(*[L._p(t) for t in a], eps, flags, L.stream_ptr(x.device)), "kon_attn_fwd")
        ctx.save_for_backward(x, wq, wk, wr, gamma, beta)
        ctx.flags, ctx.eps = flags, eps
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.lib()
        x, wq, wk, wr, gamma, beta = ctx.saved_tensors
        dev = x.device
        # the bf16 backward reads gy through its strides (a permuted window of the consumer's gradient)
        strided_ok = (ctx.flags & L.KON_ATTN_BF16) and gy.stride(-1) == 1 and x.shape[2] <= 32 and wq.shape[1] <= 4 \
            and all(s % 2 == 0 for s in gy.stride()[:3]) and gy.data_ptr() % 8 == 0
        if not strided_ok:
            gy = gy.contiguous()
        dx = torch.empty_like(x)
        dwq, dwk = torch.empty_like(wq), torch.empty_like(wk)
        dwr = None if wr is None else torch.empty_like(wr)
        dg = None if gamma is None else torch.empty_like(gamma)
        db = None if beta is None else torch.empty_like(beta)
        ws = _ws(lib.kon_attn_bwd_workspace_bytes(x.shape[0], x.shape[1], x.shape[2], wq.shape[1],
                                                  wq.shape[2], dev.index or 0), dev)
        a = [L._arg(t) for t in (x, wq, wk, wr, gamma, beta, gy, dx, dwq, dwk, dwr, dg, db, ws)]
        with _prof("attn_bwd"):
            L.check(lib.kon_attn_bwd(*[L._p(t) for t in a], ctx.eps, ctx.flags, L.stream_ptr(dev)), "kon_attn_bwd")
        return dx, dwq, dwk, dwr, dg, db, None, None, None


def attention(x, wq, wk, wr=None, gamma=None, beta=None, use_scale=True, use_ln=True, use_res=True,
              relu=True, eps: float = 1e-3, bf16: bool = False, layout: str = "hbfd"):
    """x [B,F,kin]; w* [kin,H,d] -> [H,B,F,d] = ReLU(LN(sigmoid(QK^T/sqrt d) K) + X Wr).
    ``layout`` (bf16 path): memory order of the result, see ``_ATTN_LAYOUTS``."""
    flags = ((L.KON_ATTN_USE_SCALE if use_scale else 0) | (L.KON_ATTN_USE_LN if use_ln else 0) |
             (L.KON_ATTN_USE_RES if use_res else 0) | (L.KON_ATTN_RELU if relu else 0) |
             (L.KON_ATTN_BF16 if bf16 else 0))
    if layout not in _ATTN_LAYOUTS:
        raise L.KonError(f"attention layout {layout!r} not in {sorted(_ATTN_LAYOUTS)}")
    if not bf16:
        layout = "hbfd"
    return _Attn.apply(x, wq, wk, wr if use_res else None, gamma if use_ln else None,
                       beta if use_ln else None, flags, eps, layout)


# --------------------------------------------------------------------------- #
# a11: the concat buffer (StackLayer without the copy)
# --------------------------------------------------------------------------- #
# --------------------------------------------------------------------------- #
# skinny heads (a12): y = [x1 | x2] W + b, N in {1,2}
# --------------------------------------------------------------------------- #
class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, w, b):
        lib = L.lib()
        y = torch.empty((x1.shape[0], w.shape[1]), dtype=torch.float32, device=x1.device)
        a = [L._arg(t) for t in (x1, x2, w, b, y)]
        with _prof("head_fwd"):
            L.check(lib.kon_head_fwd(*[L._p(t) for t in a], L.stream_ptr(x1.device)), "kon_head_fwd")
        ctx.save_for_backward(x1, x2, w)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.lib()
        x1, x2, w = ctx.saved_tensors
        dev = x1.device
        gy = gy.contiguous()
        dx1 = torch.empty_like(x1, memory_format=torch.contiguous_format) if ctx.needs_input_grad[0] else None
        dx2 = (torch.empty_like(x2, memory_format=torch.contiguous_format)
               if (x2 is not None and ctx.needs_input_grad[1]) else None)
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[1], dtype=torch.float32, device=dev) if ctx.has_b else None
        ws = _ws(lib.kon_head_bwd_workspace_bytes(x1.shape[0], w.shape[0], w.shape[1], dev.index or 0), dev)
        a = [L._arg(t) for t in (x1, x2, w, gy, dx1, dx2, dw, db, ws)]
        with _prof("head_bwd"):
            L.check(lib.kon_head_bwd(*[L._p(t) for t in a], L.stream_ptr(dev)), "kon_head_bwd")
        return dx1, dx2, dw, db


def head_supported(x1, x2, w) -> bool:
    d = x1.shape[1] + (0 if x2 is None else x2.shape[1])
    return (x1.dim() == 2 and x1.dtype == torch.float32 and w.dtype == torch.float32 and w.shape[1] in (1, 2)
            and d <= 1024 and x1.stride(1) == 1
            and (x2 is None or (x2.dim() == 2 and x2.dtype == torch.float32 and x1.shape[1] % 4 == 0 and x2.stride(1) == 1)))


def head(x1, x2, w, b=None):
    """``[x1 | x2] @ w + b`` for the 1- and 2-unit heads, both inputs read in place (no concat)."""
    return _Head.apply(x1, x2, w, b)


class _EmbedConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, arena, ids, field_row_offset, dense, width):
        B, F = ids.shape[0], ids.shape[1]
        dim = arena.shape[1]
        Fk = F * dim
        nd = 0 if dense is None else dense.shape[1]
        assert width % 4 == 0 and width >= Fk + nd
        if arena.requires_grad:
            embed_presort(ids, field_row_offset)
            _announce_main(arena, ids, field_row_offset)
        xcat = torch.empty((B, width), dtype=arena.dtype, device=arena.device)
        embed_fwd_raw(arena.detach(), ids, field_row_offset, False, out=xcat[:, :Fk].view(B, F, dim))
        if nd:
            xcat[:, Fk:Fk + nd] = dense
        if width > Fk + nd:
            xcat[:, Fk + nd:] = 0
        ctx.save_for_backward(ids)
        ctx.arena, ctx.offs, ctx.Fk, ctx.nd = arena, field_row_offset, Fk, nd
        ctx.dense_grad = dense is not None and dense.requires_grad
        return xcat

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        arena = ctx.arena
        B, F = ids.shape[0], ids.shape[1]
        if g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16:
            g = g.contiguous()
        if arena.requires_grad:
            _main_backward(arena, g[:, :ctx.Fk].view(B, F, ctx.Fk // F), ids, ctx.offs)
        gd = g[:, ctx.Fk:ctx.Fk + ctx.nd] if ctx.dense_grad else None
        return None, None, None, gd, None


def embed_lookup_concat(arena, ids, field_row_offset, dense, width):
    """-> xcat [B,width] = [F*dim embedding columns | dense | zero pad] (see models.py)."""
    return _EmbedConcat.apply(arena, ids, tuple(field_row_offset), dense, width)


# --------------------------------------------------------------------------- #
# callers / siblings wired from the same layer classes: sum pooling of a materialised
# sequence, the un-summed pairwise products, product attention on explicit q, k, v with masks
# --------------------------------------------------------------------------- #
class _PoolSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = L.lib()
        if x.stride(2) != 1 and x.shape[2] != 1:
            x = x.contiguous()
        out = torch.empty((x.shape[0], x.shape[2]), dtype=x.dtype, device=x.device)
        a, o = L._arg(x), L._arg(out)
        with _prof("pool_sum"):
            L.check(lib.kon_pool_sum_fwd(a.ptr, o.ptr, L.stream_ptr(x.device)), "kon_pool_sum_fwd")
        ctx.L = x.shape[1]
        return out

    @staticmethod
    def backward(ctx, g):
        return g.unsqueeze(1).expand(g.shape[0], ctx.L, g.shape[1])     # d(sum)/dx: a broadcast, no kernel


def pool_sum(x: torch.Tensor) -> torch.Tensor:
    """x [B,L,k] -> [B,k] = sum over axis 1 in order l = 0..L-1 (BL:46, IL:364)."""
    return _PoolSum.apply(x)


class _Pairs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        lib = L.lib()
        if v.stride(2) != 1 and v.shape[2] != 1:
            v = v.contiguous()
        F = v.shape[1]
        out = torch.empty((v.shape[0], F * (F - 1) // 2, v.shape[2]), dtype=v.dtype, device=v.device)
        a, o = L._arg(v), L._arg(out)
        with _prof("pairs_fwd"):
            L.check(lib.kon_pairs_fwd(a.ptr, o.ptr, L.stream_ptr(v.device)), "kon_pairs_fwd")
        ctx.save_for_backward(v)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        (v,) = ctx.saved_tensors
        g = g.contiguous()
        dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
        a, b, c = L._arg(v), L._arg(g), L._arg(dv)
        with _prof("pairs_bwd"):
            L.check(lib.kon_pairs_bwd(a.ptr, b.ptr, c.ptr, L.stream_ptr(v.device)), "kon_pairs_bwd")
        return dv


def pairs(v: torch.Tensor) -> torch.Tensor:
    """v [B,F,k] -> [B, F(F-1)/2, k]: v_i * v_j for i<j in itertools.combinations order (IL:61)."""
    return _Pairs.apply(v)


class _Pattn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, use_scale, mask_mode):
        lib = L.lib()
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out = torch.empty_like(q)
        a = [L._arg(t) for t in (q, k, v, mask, out)]
        with _prof("pattn_fwd"):
            L.check(lib.kon_pattn_fwd(*[L._p(t) for t in a], 1 if use_scale else 0, mask_mode,
                                      L.stream_ptr(q.device)), "kon_pattn_fwd")
        ctx.save_for_backward(q, k, v, mask)
        ctx.use_scale, ctx.mask_mode = use_scale, mask_mode
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        q, k, v, mask = ctx.saved_tensors
        g = g.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        a = [L._arg(t) for t in (q, k, v, mask, g, dq, dk, dv)]
        with _prof("pattn_bwd"):
            L.check(lib.kon_pattn_bwd(*[L._p(t) for t in a], 1 if ctx.use_scale else 0, ctx.mask_mode,
                                      L.stream_ptr(q.device)), "kon_pattn_bwd")
        return dq, dk, dv, None, None, None


def product_attention(q, k, v, mask=None, use_scale=False, mask_mode=0):
    """sigmoid(mask(q k^T [/ sqrt d])) v on explicit q, k, v [..., F, d] (BL:292-311).  ``mask`` [F,F]:
    mode 1 ``score @ mask``, mode 2 ``score + mask * -1e5`` (BL:299-306); boolean masks are cast to float
    exactly as the reference does (``tf.cast(mask, 'float')``)."""
    if mask is None:
        mask_mode = 0
    else:
        if mask_mode not in (1, 2):
            raise L.KonError("product_attention: mask given but mask_mode is %r (must be 1 or 2)" % (mask_mode,))
        if mask.dim() != 2:
            raise L.KonError("product_attention: only a [F,F] mask shared by all samples is provided")
        mask = mask.to(torch.float32).contiguous()
    return _Pattn.apply(q, k, v, mask, bool(use_scale), int(mask_mode))


class _Bce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, y, eps):
        lib = L.lib()
        p, y = p.contiguous(), y.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        ws = _ws(lib.kon_bce_workspace_bytes(), p.device)
        a = [L._arg(t) for t in (p, y, loss, ws)]
        with _prof("bce_fwd"):
            L.check(lib.kon_bce_fwd(*[t.ptr for t in a], eps, L.stream_ptr(p.device)), "kon_bce_fwd")
        ctx.save_for_backward(p, y)
        ctx.eps = eps
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        p, y = ctx.saved_tensors
        g = g.reshape(1).contiguous().float()
        dp = torch.empty_like(p)
        a = [L._arg(t) for t in (p, y, g, dp)]
        with _prof("bce_bwd"):
            L.check(lib.kon_bce_bwd(*[t.ptr for t in a], ctx.eps, L.stream_ptr(p.device)), "kon_bce_bwd")
        return dp, None, None


def binary_crossentropy(y_true: torch.Tensor, y_pred: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """``compile(loss=binary_crossentropy)`` on probabilities, fused: clip, log terms, mean over the last axis
    and the batch (= mean over all elements), deterministic; the gradient in one kernel."""
    if y_true.shape != y_pred.shape:
        y_true = y_true.reshape(y_pred.shape)
    return _Bce.apply(y_pred.float(), y_true.float(), float(eps))
