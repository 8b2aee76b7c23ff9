"""Multi-GPU execution of the CTR hot path: one process per GPU, ``torch.distributed`` (NCCL
over NVLink 5 / NVSwitch on the B200 box, gloo in the CPU unit tests of the routing logic).

The reference is single-process (no distributed code at all); this is the scale-out
of the same step:

* dense interaction + MLP weights are replicated, the batch is split (data parallel), dense
  gradients are summed with ONE flat-bucket ``all_reduce`` per step;
* embedding tables are sharded: small tables table-wise (whole table on one rank, greedy by
  lookup count), tables with >= ``row_wise_min_rows`` rows row-wise (row r lives on rank
  ``r % N`` at local row ``r // N``).  ids are all-gathered (4 B per lookup vs 4k B of payload);
  every owner gathers for the GLOBAL batch with the same ``kon_embed_fwd`` kernel, then

      table-wise part : ``all_to_all_single``  -> each rank gets its samples' rows
      row-wise part   : ``reduce_scatter``      (non-owned ids gather zeros, the sum assembles)

  and the backward mirrors it (``all_to_all_single`` / ``all_gather`` of dOut, then the local
  sort-then-segment ``kon_embed_bwd``; non-owned ids carry no gradient).
  The first-order (dim-1) tables only ever enter the models through their sum over fields, so
  each rank sums its own fields/rows and one tiny ``reduce_scatter`` finishes the job.

On NCCL/CUDA the payload exchange is FUSED into the embedding kernels (``PeerRegion``,
``_PeerLookup``; C-ABI ``kon_embed_fwd_peer`` / ``kon_embed_bwd_peer`` / ``kon_peer_barrier``): every
rank owns one CUDA-IPC region ``[flags | xcat | dxcat]`` mapped by all the others; the gather kernel
stores each row straight into the concat buffer of the sample's rank over NVLink, the segmented
reduction of the backward loads each gradient row straight from the rank that produced it, and a
one-CTA flag barrier replaces the collective.  Only the 4-byte ids (and the dim-1 sums) still go
through NCCL.  ``KON_PEER_EXCHANGE=0`` selects the NCCL all-to-all path above (the baseline).
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn


class ShardPlan:
    """Which rank owns which (field, row)."""

    def __init__(self, rows: Sequence[int], world: int, row_wise_min_rows: int = 50_000_000, balance: str = "count",
                 row_bytes: int = 64):
        self.rows = [int(r) for r in rows]
        self.world = world
        self.balance = balance
        F = len(self.rows)
        self.rw_fields = [f for f in range(F) if world > 1 and self.rows[f] >= row_wise_min_rows]
        self.tw_fields = [f for f in range(F) if f not in self.rw_fields]
        self.tw_owner = {}
        n_tw = len(self.tw_fields)
        if balance == "cost" and world > 1:
            # Every field costs the same number of lookups and the same NVLink bytes, but not the same HBM time: a
            # lookup into a table that stays in the 126 MB L2 is a hit, a lookup into a multi-GB table is a random
            # DRAM row (gather, scatter-add and row-wise Adam alike).  Contiguous count-balanced blocks put the two
            # largest Criteo tables on rank 0 and none on four other ranks, and everybody waits for rank 0 at the
            # barriers.  Greedy placement, heaviest first, onto the least loaded rank, with the per-rank field count
            # capped so that the NVLink egress stays balanced too.
            cap = (n_tw + world - 1) // world
            wgt = {f: 1.0 + 3.0 * min(1.0, self.rows[f] * row_bytes / 64e6) for f in self.tw_fields}
            load, cnt = [0.0] * world, [0] * world
            for f in sorted(self.tw_fields, key=lambda f_: (-wgt[f_], f_)):
                p = min((q for q in range(world) if cnt[q] < cap), key=lambda q: (load[q], cnt[q], q))
                self.tw_owner[f] = p
                load[p] += wgt[f]
                cnt[p] += 1
        else:
            # balanced by field COUNT; contiguous blocks of fields per rank keep the exchanged layout identical to
            # the model's field order (no permutation copy after an NCCL all-to-all)
            base, extra = divmod(n_tw, world)
            pos = 0
            for p in range(world):
                cnt = base + (1 if p < extra else 0)
                for f in self.tw_fields[pos:pos + cnt]:
                    self.tw_owner[f] = p
                pos += cnt
        self.tw_of_rank = [[f for f in self.tw_fields if self.tw_owner[f] == p] for p in range(world)]
        # field order after the exchange ("rank-major"): rank 0's tw fields, rank 1's, ..., then rw fields
        self.exchange_order = [f for p in range(world) for f in self.tw_of_rank[p]] + self.rw_fields
        inv = [0] * F
        for pos, f in enumerate(self.exchange_order):
            inv[f] = pos
        self.to_global = inv            # emb_global[:, f] = emb_exchange[:, to_global[f]]
        self.identity_order = self.exchange_order == list(range(F))

    def local_rows(self, rank: int, f: int) -> int:
        if f in self.rw_fields:
            return (self.rows[f] - rank + self.world - 1) // self.world
        return self.rows[f]

    def local_offsets(self, rank: int):
        """(tw offsets [n_tw+1], rw offsets [n_rw+1]) into the rank's arena (tw tables first)."""
        o = [0]
        for f in self.tw_of_rank[rank]:
            o.append(o[-1] + self.local_rows(rank, f))
        tw = list(o)
        o = [tw[-1]]
        for f in self.rw_fields:
            o.append(o[-1] + self.local_rows(rank, f))
        return tw, o

    @staticmethod
    def runs(fields: Sequence[int]):
        """Ascending field list -> [(first global field, position of it in the list, length)] of maximal runs of
        ADJACENT global fields (one strided 2-D copy each)."""
        out = []
        for j, f in enumerate(fields):
            if out and out[-1][0] + out[-1][2] == f:
                out[-1] = (out[-1][0], out[-1][1], out[-1][2] + 1)
            else:
                out.append((f, j, 1))
        return out

    def describe(self) -> str:
        return (f"dp{self.world} dense + sharded embeddings ({len(self.tw_fields)} table-wise"
                f"{' cost-balanced' if self.balance == 'cost' else ''}, {len(self.rw_fields)} row-wise fields)")


# ------------------------------------------------------------------------------------------
# collectives (NCCL on GPU; gloo lacks reduce_scatter, so the CPU tests use all_reduce + slice)
# ------------------------------------------------------------------------------------------
def _reduce_scatter(out: torch.Tensor, inp: torch.Tensor, group):
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, inp, group=group)
    else:
        t = inp.clone()
        dist.all_reduce(t, group=group)
        n = out.numel()
        out.copy_(t.reshape(-1)[dist.get_rank(group) * n:(dist.get_rank(group) + 1) * n].view_as(out))


def _all_gather(out: torch.Tensor, inp: torch.Tensor, group):
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, inp, group=group)
    else:
        parts = [torch.empty_like(inp) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, inp, group=group)
        out.copy_(torch.cat([p.reshape(-1) for p in parts]).view_as(out))


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


def _fire_backward_hook(sh):
    """The embedding lookups are the first nodes of the forward graph, hence the LAST nodes autograd runs, and
    AccumulateGrad nodes run at top priority: when an embedding backward starts, every dense-weight gradient of the
    step is final.  The trainer hooks the dense all-reduce here so that it overlaps the embedding exchange / scatter."""
    hook = getattr(sh, "backward_hook", None)
    if hook is not None:
        hook()


class _ShardedLookup(torch.autograd.Function):
    """ids_local [B_l,F] -> emb [B_l,F,k] in global field order."""

    @staticmethod
    def forward(ctx, arena, ids_local, sh: "ShardedEmbed"):
        plan, g, N, rank = sh.plan, sh.group, sh.world, sh.rank
        B_l, F = ids_local.shape
        k = arena.shape[1]
        dev = arena.device
        ids_tw, ids_rw = sh.exchange_ids(ids_local)
        n_tw, n_rw = len(plan.tw_of_rank[rank]), len(plan.rw_fields)
        chunks = []
        if len(plan.tw_fields):
            send = sh.lookup_fn(arena, ids_tw, sh.tw_offs) if n_tw else torch.empty((N * B_l, 0, k), device=dev)
            in_splits = [B_l * n_tw * k] * N
            out_splits = [B_l * len(plan.tw_of_rank[p]) * k for p in range(N)]
            recv = torch.empty(sum(out_splits), dtype=arena.dtype, device=dev)
            _all_to_all(recv, send.reshape(-1), out_splits, in_splits, g)
            o = 0
            for p in range(N):
                if out_splits[p]:
                    chunks.append(recv[o:o + out_splits[p]].view(B_l, len(plan.tw_of_rank[p]), k))
                o += out_splits[p]
        if n_rw:
            part = sh.lookup_fn(arena, ids_rw, sh.rw_offs)                   # [B_g, n_rw, k], zeros where not owned
            mine = torch.empty((B_l, n_rw, k), dtype=arena.dtype, device=dev)
            _reduce_scatter(mine, part, g)
            chunks.append(mine)
        ex = chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=1)     # exchange (rank-major) order
        emb = ex if plan.identity_order else ex.index_select(1, sh.to_global)
        ctx.sh, ctx.arena = sh, arena
        ctx.save_for_backward(ids_tw, ids_rw)
        ctx.B_l = B_l
        return emb

    @staticmethod
    def backward(ctx, gout):
        _fire_backward_hook(ctx.sh)
        sh, arena = ctx.sh, ctx.arena
        ids_tw, ids_rw = ctx.saved_tensors
        plan, g, N, rank = sh.plan, sh.group, sh.world, sh.rank
        B_l, k = ctx.B_l, arena.shape[1]
        dev = arena.device
        gex = gout if plan.identity_order else gout.index_select(1, sh.to_exchange)   # exchange order
        n_tw, n_rw = len(plan.tw_of_rank[rank]), len(plan.rw_fields)
        n_tw_all = len(plan.tw_fields)
        grads = []
        if n_tw_all:
            # slab p of the exchange order goes back to rank p
            send = torch.cat([gex[:, o:o + c].reshape(-1) for o, c in sh.tw_slabs])
            in_splits = [B_l * len(plan.tw_of_rank[p]) * k for p in range(N)]
            out_splits = [B_l * n_tw * k] * N
            recv = torch.empty(sum(out_splits), dtype=gout.dtype, device=dev)
            _all_to_all(recv, send.contiguous(), out_splits, in_splits, g)
            if n_tw:
                grads.append((recv.view(N * B_l, n_tw, k), ids_tw, sh.tw_offs))
        if n_rw:
            mine = gex[:, n_tw_all:].contiguous()
            full = torch.empty((N * B_l, n_rw, k), dtype=gout.dtype, device=dev)
            _all_gather(full, mine, g)
            grads.append((full, ids_rw, sh.rw_offs))
        if arena.requires_grad:
            if not hasattr(arena, "kon_sparse_grads"):
                arena.kon_sparse_grads = []
            first = len(arena.kon_sparse_grads) == 0
            for gg, ii, oo in grads:
                arena.kon_sparse_grads.append(sh.scatter_fn(gg, ii, oo))
                arena.kon_sparse_grads[-1].disjoint = first
        return None, None, None


# ------------------------------------------------------------------------------------------
# peer memory (CUDA IPC over NVLink / NVSwitch)
# ------------------------------------------------------------------------------------------
class _RawCuda:
    """``__cuda_array_interface__`` holder so torch can view memory the library allocated."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerRegion:
    """One exchange region per rank, mapped by every other rank.

    Layout (bytes): ``[0,256)`` barrier flag block, then 256-B aligned sub-buffers handed out by
    ``carve``.  ``ptrs[q]`` is THIS process's address of rank q's region."""

    FLAG_BYTES = 256

    def __init__(self, group, device: torch.device, nbytes: int):
        from . import _lib as L
        self.L, self.lib = L, L.lib()
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        L.check(self.lib.kon_peer_alloc(self.dev_index, self.nbytes, ctypes.byref(ptr), handle), "kon_peer_alloc")
        self.local_ptr = ptr.value
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), self.nbytes), group=group)
        self.ptrs: List[int] = []
        for q, (h, nb) in enumerate(handles):
            if nb != self.nbytes:
                raise RuntimeError(f"PeerRegion: rank {q} allocated {nb} bytes, this rank {self.nbytes}")
            if q == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            pp = ctypes.c_void_p()
            hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            L.check(self.lib.kon_peer_open(self.dev_index, hb, ctypes.byref(pp)), "kon_peer_open")
            self.ptrs.append(pp.value)
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*self.ptrs)
        self._bytes = torch.as_tensor(_RawCuda(self.local_ptr, self.nbytes), device=device)
        self._used = self.FLAG_BYTES
        self._n_barriers = 0
        self.flags = self._bytes[:self.FLAG_BYTES].view(torch.int32)
        dist.barrier(group=group)       # every mapping exists before anybody stores through one

    def carve(self, shape, dtype=torch.float32):
        """-> (local tensor view, byte offset inside the region)."""
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        off = self._used
        if off + nb > self.nbytes:
            raise RuntimeError("PeerRegion exhausted")
        self._used = (off + nb + 255) // 256 * 256
        return self._bytes[off:off + nb].view(dtype).view(*shape), off

    def ptr_array(self, byte_offset: int):
        return (ctypes.c_void_p * self.world)(*[p + byte_offset for p in self.ptrs])

    def barrier(self, timeout_ms: Optional[int] = None):
        """Flag barrier on the current stream.  A peer that does not arrive within the timeout makes the kernel set
        the region's error word instead of hanging the GPU; ``check()`` turns that into a ``KonError`` (the trainer
        calls it on a cadence, see ``train.Trainer``).  Default 10 s (``KON_PEER_TIMEOUT_MS``), 60 s for the first
        barriers of a region (lazy initialisation / graph warm-up skew between the ranks)."""
        if timeout_ms is None:
            timeout_ms = int(os.environ.get("KON_PEER_TIMEOUT_MS", "10000"))
            if self._n_barriers < 8:
                timeout_ms = max(timeout_ms, 60000)
        self._n_barriers += 1
        from . import ops
        with ops._prof("peer_barrier"):                   # device time of the barrier = flag round trip + waiting for the slowest rank
            self.L.check(self.lib.kon_peer_barrier(self._flag_ptrs, self.world, self.rank, self.dev_index, timeout_ms,
                                                   self.L.stream_ptr(self.device)), "kon_peer_barrier")

    def check(self):
        """Raise if a barrier of this region ever timed out (synchronises the device)."""
        if self.timed_out():
            raise self.L.KonError(
                f"rank {self.rank}: a peer barrier timed out -- the embedding rows / gradients exchanged in that step "
                "were incomplete; the training state is not trustworthy from that step on")

    def timed_out(self) -> bool:
        """True when a barrier gave up waiting for a peer (word 17 of the flag block); synchronises."""
        return bool(self.flags[17].item() != 0)

    def close(self):
        if getattr(self, "local_ptr", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody unmaps while a peer may still touch the region
        for q, p in enumerate(self.ptrs):
            if q != self.rank:
                self.lib.kon_peer_close(self.dev_index, ctypes.c_void_p(p))
        dist.barrier(group=self.group)
        self._bytes = self.flags = None
        self.lib.kon_peer_free(self.dev_index, ctypes.c_void_p(self.local_ptr))
        self.local_ptr = None


def _put2d(sh, puts):
    """puts: list of (src_ptr, dst_ptr, src_pitch, dst_pitch, width, rows) in bytes -> one kon_peer_put2d launch."""
    from . import _lib as L
    puts = [p for p in puts if p[4] > 0 and p[5] > 0]
    if not puts:
        return
    arr = (L.KonPut2D * len(puts))(*[L.KonPut2D(*p) for p in puts])
    dev = sh.arena.device
    from . import ops
    with ops._prof("peer_put"):
        L.check(L.lib().kon_peer_put2d(arr, len(puts), dev.index if dev.index is not None else torch.cuda.current_device(),
                                       L.stream_ptr(dev)), "kon_peer_put2d")


class _PeerLookup(torch.autograd.Function):
    """ids_local [B_l,F] (+ dense [B_l,nd]) -> xcat [B_l,width] = [emb (F*k, global field order) | dense | 0].

    Everything the sharded embedding step exchanges travels as peer STORES over NVLink, published by flag barriers
    (three per step) -- no NCCL call on this path:
      forward   (A) the 4-byte ids: every rank stores each owner's columns into the owner's region   [kon_peer_put2d]
                    -- barrier A --
                (1) the owners gather for the GLOBAL batch and store every row straight into the concat buffer of
                    the sample's rank [kon_embed_fwd_peer]; the first-order partial sums of the linear partner
                    layer ride along [kon_peer_put2d]
                    -- barrier 1 --
      backward  (2) every rank stores each owner's columns of its output gradient into the owner's receive buffer
                    [kon_peer_put2d] (stores are fire-and-forget: the 95 MB/GPU move at NVLink bandwidth, where the
                    r1 design's row loads from the scatter kernel ran at NVLink latency); the linear partner's
                    gradient rides along
                    -- barrier 2 --
                    the owner's sort-then-segment scatter-add reads LOCAL memory only [kon_embed_bwd].
    Buffer reuse across steps is ordered by barrier A of the next step (a rank reaches it only after its backward)."""

    @staticmethod
    def forward(ctx, arena, ids_local, dense, sh: "ShardedEmbed", width: int):
        from . import ops
        plan, N, rank = sh.plan, sh.world, sh.rank
        B_l, F = ids_local.shape
        k = arena.shape[1]
        px = sh.peer_buffers(B_l, width)
        ids_tw, ids_rw = sh.exchange_ids(ids_local, px)  # (A) + barrier A
        xcat, region = px["xcat"], px["region"]
        n_tw, n_rw, n_tw_all = len(plan.tw_of_rank[rank]), len(plan.rw_fields), len(plan.tw_fields)
        if arena.requires_grad:                           # the backward's routing: on the side stream, from here on
            if n_tw:
                ops.embed_presort(ids_tw, sh.tw_offs)
            if n_rw:
                ops.embed_presort(ids_rw, sh.rw_offs)
        base = region.ptr_array(px["xcat_off"])          # every field lands at ITS column of the model's order
        if n_tw:
            ops.embed_fwd_peer(arena.detach(), ids_tw, sh.tw_offs, base, N, B_l, width, k,
                               field_col=[f * k for f in plan.tw_of_rank[rank]])
        if n_rw:
            ops.embed_fwd_peer(arena.detach(), ids_rw, sh.rw_offs, base, N, B_l, width, k, skip_invalid=True,
                               field_col=[f * k for f in plan.rw_fields])
        nd = 0
        if dense is not None:
            nd = dense.shape[1]
            xcat[:, F * k:F * k + nd].copy_(dense)
        if width > F * k + nd:
            xcat[:, F * k + nd:].zero_()
        # the linear partner's first-order partial sums ride on the same barrier
        lp = sh.lin_partner
        lin_ids = None
        if lp is not None and lp._seen_use and ops._SHARE_SORT:
            lin_ids = ids_tw if n_rw == 0 else torch.cat([ids_tw, ids_rw], dim=1).contiguous()
            part = lp.lookup_fn(lp.arena, lin_ids, lp.all_offs, True)             # [B_g, 1]
            # one contiguous B_l-float block per (source rank, sample owner): 16-byte NVLink stores, not B_l 4-byte ones
            _put2d(sh, [(part.data_ptr() + q * B_l * 4, region.ptrs[q] + px["linparts_off"] + rank * B_l * 4,
                         B_l * 4, B_l * 4, B_l * 4, 1) for q in range(N)])
        region.barrier()                                  # barrier 1: rows (and partial sums) complete
        if lin_ids is not None:
            ops._STEP_CACHE[("lin_fwd", id(plan))] = (px["linparts"].sum(dim=0).unsqueeze(1), lin_ids, px)
        out = xcat
        ctx.sh, ctx.arena, ctx.px = sh, arena, px
        ctx.save_for_backward(ids_tw, ids_rw)
        ctx.dims = (B_l, F, k, nd, width)
        # a fresh tensor object every call (autograd attaches this call's history to it); the storage is
        # the persistent exchange buffer: ONE step in flight, like the static buffers of a CUDA graph
        return out.view(B_l, width)

    @staticmethod
    def backward(ctx, gout):
        _fire_backward_hook(ctx.sh)
        from . import ops
        sh, arena, px = ctx.sh, ctx.arena, ctx.px
        ids_tw, ids_rw = ctx.saved_tensors
        plan, N, rank = sh.plan, sh.world, sh.rank
        B_l, F, k, nd, width = ctx.dims
        region = px["region"]
        n_tw, n_rw, n_tw_all = len(plan.tw_of_rank[rank]), len(plan.rw_fields), len(plan.tw_fields)
        if gout.stride(1) == 1 and gout.stride(0) % 4 == 0 and gout.data_ptr() % 16 == 0:
            gsrc, pitch = gout, gout.stride(0) * 4       # columns [0, F*k) of the concat-buffer gradient, in place
        else:
            gsrc, pitch = gout[:, :F * k].contiguous(), F * k * 4
        puts = []
        for q in range(N):                                # owner q's columns -> rows [rank*B_l, ...) of its receive buffer
            c = len(plan.tw_of_rank[q])
            for f0, j0, cnt in sh.runs_tw[q]:
                puts.append((gsrc.data_ptr() + f0 * k * 4,
                             region.ptrs[q] + px["drecv_tw_off"] + (rank * B_l * c + j0) * k * 4,
                             pitch, c * k * 4, cnt * k * 4, B_l))
        if n_rw:
            # row-wise columns: EVERY rank needs them (its rows of the tables).  Scattered over the row they would go
            # out as N x (#runs) narrow stores; packed once locally they go out as one wide row per sample and peer.
            if len(sh.runs_rw) > 1:
                stage = torch.empty((B_l, n_rw * k), dtype=gout.dtype, device=gout.device)
                _put2d(sh, [(gsrc.data_ptr() + f0 * k * 4, stage.data_ptr() + j0 * k * 4, pitch, n_rw * k * 4, cnt * k * 4, B_l)
                            for f0, j0, cnt in sh.runs_rw])
                rsrc, rpitch = stage.data_ptr(), n_rw * k * 4
            else:
                rsrc, rpitch = gsrc.data_ptr() + sh.runs_rw[0][0] * k * 4, pitch
            for q in range(N):
                puts.append((rsrc, region.ptrs[q] + px["drecv_rw_off"] + rank * B_l * n_rw * k * 4, rpitch, n_rw * k * 4,
                             n_rw * k * 4, B_l))
        for i in range(0, len(puts), 64):
            _put2d(sh, puts[i:i + 64])
        region.barrier()                                  # barrier 2: every rank's dOut (and first-order gradient) is in place
        if arena.requires_grad:
            if not hasattr(arena, "kon_sparse_grads"):
                arena.kon_sparse_grads = []
            first = len(arena.kon_sparse_grads) == 0          # table-wise and row-wise fields: disjoint row ranges
            if n_tw:
                d_tw = px["drecv_tw"][:N * B_l * n_tw * k].view(N * B_l, n_tw, k)
                sg = None
                if n_rw == 0 and ops.FUSE_LIN:            # the first-order gradient rides the same segmented reduce
                    deferred = ops._STEP_CACHE.pop(("lin_bwd", id(plan)), None)
                    if deferred is not None:
                        sg = deferred(main=(d_tw, ids_tw, sh.tw_offs))
                if sg is None:
                    sg = ops.embed_bwd_raw(d_tw, ids_tw, sh.tw_offs)
                arena.kon_sparse_grads.append(sg)
                arena.kon_sparse_grads[-1].disjoint = first
            if n_rw:
                d_rw = px["drecv_rw"].view(N * B_l, n_rw, k)
                arena.kon_sparse_grads.append(ops.embed_bwd_raw(d_rw, ids_rw, sh.rw_offs))
                arena.kon_sparse_grads[-1].disjoint = first
        deferred = ops._STEP_CACHE.pop(("lin_bwd", id(plan)), None)
        if deferred is not None:
            deferred()
        gdense = gout[:, F * k:F * k + nd] if nd else None
        return None, None, gdense, None, None


class _ShardedSumPeer(torch.autograd.Function):
    """First-order sum whose exchange rode on the embedding layer's barriers (see ``_PeerLookup``): the forward value
    was assembled there; the backward stores this rank's gradient into every owner's receive buffer and leaves the
    owner-side scatter-add to run right after barrier 2 of the embedding backward (same step, same stream)."""

    @staticmethod
    def forward(ctx, arena, mine, lin_ids, sh: "ShardedEmbed", px):
        ctx.sh, ctx.arena, ctx.px = sh, arena, px
        ctx.save_for_backward(lin_ids)
        return mine.clone()

    @staticmethod
    def backward(ctx, gout):
        _fire_backward_hook(ctx.sh)
        from . import ops
        sh, arena, px = ctx.sh, ctx.arena, ctx.px
        (lin_ids,) = ctx.saved_tensors
        N, rank = sh.world, sh.rank
        B_l = gout.shape[0]
        g = gout.contiguous()
        region = px["region"]
        sp = sh.sparse_partner
        _put2d(sp, [(g.data_ptr(), region.ptrs[q] + px["lin_grecv_off"] + rank * B_l * 4, B_l * 4, B_l * 4, B_l * 4, 1)
                    for q in range(N)])

        def scatter(need_barrier=False, main=None):
            """``main`` = (d_out, ids, offsets) of the embedding tables' scatter-add about to run on the same routing:
            both gradients are then reduced in one pass and the embedding tables' SparseGrad is returned."""
            if need_barrier:                              # no embedding backward followed in this step (ops.end_step)
                region.barrier()
            if not arena.requires_grad:
                return None
            full = px["lin_grecv"].view(N * B_l, 1)
            g3 = full.unsqueeze(1).expand(N * B_l, lin_ids.shape[1], 1)
            if not hasattr(arena, "kon_sparse_grads"):
                arena.kon_sparse_grads = []
            if main is not None:
                d_m, ids_m, offs_m = main
                if ids_m.data_ptr() == lin_ids.data_ptr() and ids_m.shape == lin_ids.shape \
                        and tuple(offs_m) == tuple(sh.all_offs) and sh.scatter_fn is _kon_scatter:
                    sg, sg1 = ops.embed_bwd_raw(d_m, ids_m, offs_m, lin=g3)
                    arena.kon_sparse_grads.append(sg1)
                    return sg
            arena.kon_sparse_grads.append(sh.scatter_fn(g3, lin_ids, sh.all_offs))
            return None
        ops._STEP_CACHE[("lin_bwd", id(sh.plan))] = scatter
        return None, None, None, None, None


class _ShardedSum(torch.autograd.Function):
    """ids_local [B_l,F] -> sum over fields of the dim-1 (first-order) tables, [B_l, dim]."""

    @staticmethod
    def forward(ctx, arena, ids_local, sh: "ShardedEmbed"):
        g, N = sh.group, sh.world
        B_l, F = ids_local.shape
        dev = arena.device
        ids_tw, ids_rw = sh.exchange_ids(ids_local)
        # no row-wise fields: the very tensor the embedding lookup routes with, so that the two backward
        # passes of the step share one sort (ops._SORT_CACHE)
        ids_loc = ids_tw if ids_rw.shape[1] == 0 else torch.cat([ids_tw, ids_rw], dim=1).contiguous()
        part = sh.lookup_fn(arena, ids_loc, sh.all_offs, True)                # [B_g, dim]
        mine = torch.empty((B_l, arena.shape[1]), dtype=arena.dtype, device=dev)
        _reduce_scatter(mine, part, g)
        ctx.sh, ctx.arena, ctx.B_l = sh, arena, B_l
        ctx.save_for_backward(ids_loc)
        return mine

    @staticmethod
    def backward(ctx, gout):
        _fire_backward_hook(ctx.sh)
        sh, arena = ctx.sh, ctx.arena
        (ids_loc,) = ctx.saved_tensors
        N = sh.world
        full = torch.empty((N * ctx.B_l, gout.shape[1]), dtype=gout.dtype, device=gout.device)
        _all_gather(full, gout.contiguous(), sh.group)
        if arena.requires_grad:
            g3 = full.unsqueeze(1).expand(full.shape[0], ids_loc.shape[1], full.shape[1])
            if not hasattr(arena, "kon_sparse_grads"):
                arena.kon_sparse_grads = []
            arena.kon_sparse_grads.append(sh.scatter_fn(g3, ids_loc, sh.all_offs))
        return None, None, None


def _kon_lookup(arena, ids, offs, sum_fields=False):
    from . import ops
    return ops.embed_fwd_raw(arena.detach(), ids, offs, sum_fields)


def _kon_scatter(g, ids, offs):
    from . import ops
    if g.stride(-1) != 1 or g.stride(0) % 4 or (g.stride(1) % 4 and g.stride(1) != 0):
        g = g.contiguous()
    return ops.embed_bwd_raw(g, ids, offs)


class ShardedEmbed(nn.Module):
    """Drop-in for ``layers.SparseEmbed`` on one rank of a sharded job (same ``lookup`` /
    ``lookup_concat`` / ``lookup_sum`` surface; ``arena`` holds this rank's tables and shards)."""

    def __init__(self, sparse_info: list, plan: ShardPlan, group, device, is_linear=False, seed=2020,
                 lookup_fn: Callable = _kon_lookup, scatter_fn: Callable = _kon_scatter):
        super().__init__()
        self.plan, self.group = plan, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.is_linear = is_linear
        self.lookup_fn, self.scatter_fn = lookup_fn, scatter_fn
        dims = {(i.linear_unit if is_linear else i.cross_unit) for i in sparse_info}
        assert len(dims) == 1
        self.dim = dims.pop()
        self.emb_reg = 0.0 if is_linear else float(sparse_info[0].emb_reg or 0.0)
        tw, rw = plan.local_offsets(self.rank)
        self.tw_offs, self.rw_offs = tuple(tw), tuple(rw)
        self.all_offs = tuple(tw + rw[1:])
        self.field_row_offset = self.all_offs
        n_local = rw[-1]
        arena = torch.empty(max(n_local, 1), self.dim, device=device, dtype=torch.float32)
        gd = torch.Generator(device=device).manual_seed(seed + 7919 * self.rank)
        fields = plan.tw_of_rank[self.rank] + plan.rw_fields
        for j, f in enumerate(fields):
            lim = 0.05 if is_linear else (6.0 / (plan.rows[f] + self.dim)) ** 0.5
            arena[self.all_offs[j]:self.all_offs[j + 1]].uniform_(-lim, lim, generator=gd)
        self.arena = nn.Parameter(arena)
        self.register_buffer("to_global", torch.tensor(plan.to_global, dtype=torch.long, device=device), persistent=False)
        self.register_buffer("to_exchange", torch.tensor(plan.exchange_order, dtype=torch.long, device=device), persistent=False)
        self.register_buffer("tw_idx", torch.tensor(plan.tw_of_rank[self.rank], dtype=torch.long, device=device), persistent=False)
        self.register_buffer("rw_idx", torch.tensor(plan.rw_fields, dtype=torch.long, device=device), persistent=False)
        o, slabs = 0, []
        for p in range(self.world):
            slabs.append((o, len(plan.tw_of_rank[p])))
            o += len(plan.tw_of_rank[p])
        self.tw_slabs = slabs
        self.runs_tw = [ShardPlan.runs(plan.tw_of_rank[p]) for p in range(self.world)]
        self.runs_rw = ShardPlan.runs(plan.rw_fields)
        self.tw_idx_of = [torch.tensor(plan.tw_of_rank[p], dtype=torch.long, device=device) for p in range(self.world)]
        # fused NVLink exchange: NCCL process group + CUDA + vector-width rows (KON_PEER_EXCHANGE=0: NCCL baseline)
        self.use_peer = (not is_linear and self.world > 1 and torch.device(device).type == "cuda"
                         and dist.get_backend(group) == "nccl" and self.dim % 4 == 0 and self.world <= 16
                         and os.environ.get("KON_PEER_EXCHANGE", "1") != "0")
        self._peer = {}
        # the first-order (linear) layer of the same model rides on the embedding layer's barriers (DistContext.attach
        # links the two); `_seen_use` is set by the linear layer's first lookup_sum, so models that never read the
        # first-order tables (DCN, AutoInt) do not pay for their partial sums
        object.__setattr__(self, "lin_partner", None)
        object.__setattr__(self, "sparse_partner", None)
        self._seen_use = False

    def peer_buffers(self, B_l: int, width: int):
        """The rank's exchange region for one (local batch, row width); built collectively on first use
        (every rank sees the same shapes in the same order).  Sub-buffers (same offsets on every rank; sizes use the
        largest per-rank field count): ``xcat`` [B_l,width] | ``ids_tw`` int32 [N*B_l*n_tw_max] | ``ids_rw`` int32
        [N*B_l*n_rw] | ``drecv_tw`` [N*B_l*n_tw_max*k] | ``drecv_rw`` [N*B_l*n_rw*k] | ``linparts`` [N,B_l] |
        ``lin_grecv`` [N*B_l]."""
        F, k, N = len(self.plan.rows), self.dim, self.world
        if (B_l, width) in self._peer:
            return self._peer[(B_l, width)]
        n_tw_max = max(max(len(t) for t in self.plan.tw_of_rank), 1)
        n_rw = len(self.plan.rw_fields)
        al = lambda nb: (nb + 255) // 256 * 256
        sizes = [B_l * width * 4, N * B_l * n_tw_max * 4, N * B_l * max(n_rw, 1) * 4, N * B_l * n_tw_max * k * 4,
                 N * B_l * max(n_rw, 1) * k * 4, B_l * N * 4, N * B_l * 4]
        region = PeerRegion(self.group, self.arena.device, PeerRegion.FLAG_BYTES + sum(al(x) for x in sizes))
        px = dict(region=region)
        px["xcat"], px["xcat_off"] = region.carve((B_l, width))
        px["ids_tw"], px["ids_tw_off"] = region.carve((N * B_l * n_tw_max,), torch.int32)
        px["ids_rw"], px["ids_rw_off"] = region.carve((N * B_l * max(n_rw, 1),), torch.int32)
        px["drecv_tw"], px["drecv_tw_off"] = region.carve((N * B_l * n_tw_max * k,))
        px["drecv_rw"], px["drecv_rw_off"] = region.carve((N * B_l * max(n_rw, 1) * k,))
        px["linparts"], px["linparts_off"] = region.carve((N, B_l))
        px["lin_grecv"], px["lin_grecv_off"] = region.carve((N * B_l,))
        self._peer[(B_l, width)] = px
        return px

    def check_peer(self):
        """Raise ``KonError`` if any exchange barrier of this layer timed out (synchronises)."""
        for px in self._peer.values():
            px["region"].check()

    def close_peer(self):
        """Unmap / free the exchange regions (collective; call before destroying the process group)."""
        for px in self._peer.values():
            px["region"].close()
        self._peer = {}

    def _exchange_ids_peer(self, ids_local: torch.Tensor, px):
        """The ids all-to-all as peer stores + one flag barrier (int32 ids, identity field order)."""
        N, plan, rank = self.world, self.plan, self.rank
        B_l, F = ids_local.shape
        region = px["region"]
        n_loc, n_rw, n_tw_all = len(plan.tw_of_rank[rank]), len(plan.rw_fields), len(plan.tw_fields)
        base, pitch = ids_local.data_ptr(), F * 4
        puts = []
        for q in range(N):
            c = len(plan.tw_of_rank[q])
            for f0, j0, cnt in self.runs_tw[q]:
                puts.append((base + f0 * 4, region.ptrs[q] + px["ids_tw_off"] + (rank * B_l * c + j0) * 4, pitch, c * 4,
                             cnt * 4, B_l))
            for f0, j0, cnt in self.runs_rw:
                puts.append((base + f0 * 4, region.ptrs[q] + px["ids_rw_off"] + (rank * B_l * n_rw + j0) * 4, pitch,
                             n_rw * 4, cnt * 4, B_l))
        for i in range(0, len(puts), 64):
            _put2d(self, puts[i:i + 64])
        region.barrier()                                  # barrier A (also: every rank has finished its previous step)
        # copies out of the region: the ids are read again by the backward's routing sort, after peers may have
        # stored the NEXT step's ids
        ids_tw = px["ids_tw"][:N * B_l * n_loc].view(N * B_l, n_loc).clone()
        if n_rw:
            r = px["ids_rw"][:N * B_l * n_rw].view(N * B_l, n_rw)
            own = (r % N) == rank
            ids_rw = torch.where(own, torch.div(r, N, rounding_mode="floor"), torch.full_like(r, -1)).contiguous()
        else:
            ids_rw = torch.empty((N * B_l, 0), dtype=ids_local.dtype, device=ids_local.device)
        return ids_tw, ids_rw

    def exchange_ids(self, ids_local: torch.Tensor, px=None):
        """Local ids ``[B_l,F]`` -> this rank's lookups for the GLOBAL batch:
        (table-wise ids ``[B_g, n_tw_loc]``, row-wise local row ids ``[B_g, n_rw]`` with -1 where the
        row lives on another rank).  Each owner only receives the columns it owns (all-to-all of
        4 B ids), the row-wise columns are all-gathered."""
        from . import ops
        N, g, plan = self.world, self.group, self.plan
        # the embedding and the first-order tables of a model are looked up with the same ids: inside a
        # training step (ops.new_step() ... end_step()) the second layer reuses the first one's exchange
        ckey = ("ids", id(plan), ids_local.data_ptr(), ids_local._version, tuple(ids_local.shape))
        if ops._SHARE_SORT and ckey in ops._STEP_CACHE:
            return ops._STEP_CACHE[ckey]
        if px is not None and ids_local.dtype == torch.int32 and ids_local.is_contiguous():
            out = self._exchange_ids_peer(ids_local, px)
            if ops._SHARE_SORT:
                ops._STEP_CACHE[ckey] = out
            return out
        B_l = ids_local.shape[0]
        dev = ids_local.device
        n_loc = len(plan.tw_of_rank[self.rank])
        if len(plan.tw_fields):
            send = torch.cat([ids_local[:, o:o + c].reshape(-1) if plan.identity_order else
                              ids_local.index_select(1, self.tw_idx_of[p]).reshape(-1)
                              for p, (o, c) in enumerate(self.tw_slabs)])
            recv = torch.empty(N * B_l * n_loc, dtype=ids_local.dtype, device=dev)
            _all_to_all(recv, send, [B_l * n_loc] * N, [B_l * c for _, c in self.tw_slabs], g)
            ids_tw = recv.view(N * B_l, n_loc)
        else:
            ids_tw = torch.empty((N * B_l, 0), dtype=ids_local.dtype, device=dev)
        n_rw = len(plan.rw_fields)
        if n_rw:
            mine = ids_local.index_select(1, self.rw_idx).contiguous()
            r = torch.empty((N * B_l, n_rw), dtype=ids_local.dtype, device=dev)
            _all_gather(r, mine, g)
            own = (r % N) == self.rank
            ids_rw = torch.where(own, torch.div(r, N, rounding_mode="floor"), torch.full_like(r, -1)).contiguous()
        else:
            ids_rw = torch.empty((N * B_l, 0), dtype=ids_local.dtype, device=dev)
        if ops._SHARE_SORT:
            ops._STEP_CACHE[ckey] = (ids_tw, ids_rw)
        return ids_tw, ids_rw

    def load_global_tables(self, tables: Sequence[torch.Tensor]):
        """Scatter full (global) tables into this rank's shard (tests / checkpoint import)."""
        with torch.no_grad():
            fields = self.plan.tw_of_rank[self.rank] + self.plan.rw_fields
            for j, f in enumerate(fields):
                t = tables[f].to(self.arena.device)
                if f in self.plan.rw_fields:
                    t = t[self.rank::self.world]
                self.arena[self.all_offs[j]:self.all_offs[j + 1]].copy_(t)

    def load_global_table(self, f: int, table: torch.Tensor):
        """One field's full (global) table -> this rank's shard of it (no-op when another rank owns field f)."""
        fields = self.plan.tw_of_rank[self.rank] + self.plan.rw_fields
        if f not in fields:
            return
        j = fields.index(f)
        with torch.no_grad():
            t = table.to(self.arena.device)
            if f in self.plan.rw_fields:
                t = t[self.rank::self.world]
            self.arena[self.all_offs[j]:self.all_offs[j + 1]].copy_(t)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        if self.use_peer:
            B, F = ids.shape
            return _PeerLookup.apply(self.arena, ids, None, self, F * self.dim).view(B, F, self.dim)
        return _ShardedLookup.apply(self.arena, ids, self)

    def lookup_sum(self, ids: torch.Tensor) -> torch.Tensor:
        from . import ops
        self._seen_use = True
        cached = ops._STEP_CACHE.pop(("lin_fwd", id(self.plan)), None) if ops._SHARE_SORT else None
        if cached is not None and self.sparse_partner is not None:
            mine, lin_ids, px = cached
            return _ShardedSumPeer.apply(self.arena, mine, lin_ids, self, px)
        return _ShardedSum.apply(self.arena, ids, self)

    def lookup_concat(self, ids: torch.Tensor, dense: Optional[torch.Tensor], width: int) -> torch.Tensor:
        if self.use_peer:
            return _PeerLookup.apply(self.arena, ids, dense, self, width)
        emb = _ShardedLookup.apply(self.arena, ids, self)
        B, F, k = emb.shape
        parts = [emb.reshape(B, F * k)]
        if dense is not None:
            parts.append(dense)
        pad = width - F * k - (0 if dense is None else dense.shape[1])
        if pad:
            parts.append(torch.zeros((B, pad), dtype=emb.dtype, device=emb.device))
        return torch.cat(parts, dim=1)


class DistContext:
    def __init__(self, group, device, row_wise_min_rows: int = 50_000_000, balance: Optional[str] = None):
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.row_wise_min_rows = row_wise_min_rows
        # cost-balanced placement needs the peer path (fields of an owner are not adjacent); the NCCL baseline keeps
        # contiguous blocks (KON_SHARD_BALANCE=count|cost overrides)
        # "count" (default): contiguous blocks of fields per rank -> every peer store / put moves 192-256 contiguous
        # bytes per sample.  "cost" (KON_SHARD_BALANCE=cost): one multi-GB table per rank; measured on 8xB200 it removes
        # the HBM skew (rank-0 gather 0.22 -> 0.16 ms, scatter 0.25 -> 0.18, Adam 0.10 -> 0.06) but scatters every
        # owner's fields over the row, so the NVLink writes shrink to 64-byte runs and the step does not get faster
        # (8.73 vs 8.86 ms): kept as an option, not the default.
        self.balance = balance or os.environ.get("KON_SHARD_BALANCE") or "count"
        self.plan: Optional[ShardPlan] = None

    def attach(self, model):
        """Replace the model's replicated embedding layers by sharded ones (in place)."""
        info = model.sparse_embed.sparse_info
        dim = info[0].cross_unit
        self.plan = ShardPlan([i.word_size for i in info], self.world, self.row_wise_min_rows, balance=self.balance,
                              row_bytes=4 * int(dim))
        dev = self.device
        old = model.sparse_embed
        model.sparse_embed = ShardedEmbed(info, self.plan, self.group, dev, is_linear=False, seed=old.seed)
        if model.linear_embed is not None:
            model.linear_embed = ShardedEmbed(info, self.plan, self.group, dev, is_linear=True, seed=old.seed)
            if model.sparse_embed.use_peer:      # plain attributes (not sub-modules: the link is cyclic)
                object.__setattr__(model.sparse_embed, "lin_partner", model.linear_embed)
                object.__setattr__(model.linear_embed, "sparse_partner", model.sparse_embed)
        del old
        return model

    def allreduce_dense_grads(self, params: List[torch.nn.Parameter]):
        gs = [p.grad for p in params if p.grad is not None]
        if not gs:
            return
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat, group=self.group)
        o = 0
        for g in gs:
            n = g.numel()
            g.copy_(flat[o:o + n].view_as(g))
            o += n

    # ---- the same all-reduce, overlapped with the embedding backward ---------------------------------
    def arm_overlapped_allreduce(self, model, params_fn):
        """Before ``backward()``: the first embedding backward of the step launches the dense all-reduce on a
        side stream (see ``_fire_backward_hook``); ``finish_allreduce`` joins it (or runs it in line when no hook fired)."""
        self._ar_done = False
        self._ar_event = None
        self._ar_keep = None

        def hook():
            if self._ar_done:
                return
            self._ar_done = True
            if torch.device(self.device).type != "cuda":
                self.allreduce_dense_grads(params_fn())
                return
            cur = torch.cuda.current_stream(self.device)
            if getattr(self, "_ar_stream", None) is None:
                self._ar_stream = torch.cuda.Stream(device=self.device)
            side = self._ar_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gs = [p.grad for p in params_fn() if p.grad is not None]
                if gs:
                    flat = torch.cat([g.reshape(-1) for g in gs])
                    dist.all_reduce(flat, group=self.group)
                    o = 0
                    for g in gs:
                        n = g.numel()
                        g.copy_(flat[o:o + n].view_as(g))
                        o += n
                    self._ar_keep = (flat, gs)          # alive until the main stream has joined
                ev = torch.cuda.Event()
                ev.record(side)
            self._ar_event = ev
        for emb in (model.sparse_embed, model.linear_embed):
            if emb is not None and hasattr(emb, "plan"):
                emb.backward_hook = hook

    def finish_allreduce(self, model, params):
        for emb in (model.sparse_embed, model.linear_embed):
            if emb is not None and hasattr(emb, "plan"):
                emb.backward_hook = None
        if not getattr(self, "_ar_done", False):
            self.allreduce_dense_grads(params)          # no sharded lookup took part in this backward
        elif self._ar_event is not None:
            from . import ops
            with ops._prof("allreduce_wait"):             # what the overlapped dense all-reduce still costs the main stream
                torch.cuda.current_stream(self.device).wait_event(self._ar_event)
        self._ar_keep = None

    def describe(self) -> str:
        return self.plan.describe() if self.plan else f"dp{self.world}"
