"""One training step of a CTR model the way ``model.compile(loss=binary_crossentropy,
optimizer='adam'); model.fit(...)`` runs it in the reference (example/ctr_example/
un_seq.py:61-62): forward, loss, backward, optimizer.

Dense parameters: ``torch.optim.Adam(fused=True)``.  Embedding arenas: the sparse gradient
``(unique rows, summed grads)`` from ``kon_embed_bwd`` goes straight into the row-wise
(lazy) Adam kernel ``kon_embed_adam`` together with the reference's L2 term (IL:217) -- the
dense ``[R,dim]`` gradient Keras materialises is never formed.

Multi-GPU (one process per GPU, see ``parallel.py``): dense grads are all-reduced in one flat
bucket; the embedding exchange is handled inside the sharded embedding layer.
"""
from __future__ import annotations

from typing import Optional

import os

import torch

from . import ops
from .models import XDeepFM, keras_binary_crossentropy


class SparseAdam:
    """Row-wise lazy Adam state for one arena.  The step counter lives on the device so that the
    whole training step can be captured in a CUDA graph and replayed."""

    def __init__(self, arena: torch.nn.Parameter, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, l2=0.0):
        self.arena = arena
        self.m = torch.zeros_like(arena.data)
        self.v = torch.zeros_like(arena.data)
        self.lr, self.beta1, self.beta2, self.eps, self.l2 = lr, beta1, beta2, eps, l2
        self.t = torch.zeros(1, dtype=torch.int32, device=arena.device)

    def step(self):
        sgs = getattr(self.arena, "kon_sparse_grads", None)
        if not sgs:
            return
        self.t += 1
        if len(sgs) > 1 and not all(getattr(sg, "disjoint", False) for sg in sgs):
            # the arena was looked up more than once in this step: Adam must see the SUM of the gradients of a row
            # once, not one update per lookup (rare path: merged with torch ops; needs one host sync per gradient)
            rows = torch.cat([sg.rows[:int(sg.n)] for sg in sgs]).long()
            grads = torch.cat([sg.grads[:int(sg.n)] for sg in sgs])
            uniq, inv = torch.unique(rows, return_inverse=True)
            merged = torch.zeros((uniq.numel(), grads.shape[1]), dtype=grads.dtype, device=grads.device).index_add_(0, inv, grads)
            sgs = [ops.SparseGrad(uniq.to(torch.int32), merged, torch.tensor([uniq.numel()], dtype=torch.int32, device=grads.device))]
        for sg in sgs:
            ops.embed_adam_devstep(self.arena.data, self.m, self.v, sg, self.lr, self.beta1, self.beta2,
                                   self.eps, self.l2, self.t)
        self.arena.kon_sparse_grads = []


class KerasAdam:
    """``compile(optimizer='adam')`` exactly as Keras 2.x applies it -- the reference's optimizer, for parity runs:

        m <- b1 m + (1-b1) g,  v <- b2 v + (1-b2) g^2,  w <- w - lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)

    on EVERY element of every weight at every step (Keras turns the IndexedSlices gradient of an Embedding into a dense
    update: rows a step does not touch still get the m / v decay, the dense ``2 l2 w`` regulariser term (IL:217) and a
    weight update).  For an embedding arena that is three full passes over the table per step, which is why the
    default is the row-wise lazy kernel (``SparseAdam``); trained weights of the two modes drift apart on rows that
    are not touched every step.  torch ops on the device (not CUDA-graph capturable: the step count lives on the host)."""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.state = {}
        self.t = 0

    def begin_step(self):
        self.t += 1

    def apply(self, w: torch.Tensor, g: torch.Tensor):
        key = (w.data_ptr(), tuple(w.shape))        # `p.data` is a fresh Python object on every access
        st = self.state.get(key)
        if st is None:
            st = self.state[key] = (torch.zeros_like(w), torch.zeros_like(w))
        m, v = st
        m.mul_(self.beta1).add_(g, alpha=1.0 - self.beta1)
        v.mul_(self.beta2).addcmul_(g, g, value=1.0 - self.beta2)
        lr_t = self.lr * (1.0 - self.beta2 ** self.t) ** 0.5 / (1.0 - self.beta1 ** self.t)
        w.addcdiv_(m, v.sqrt().add_(self.eps), value=-lr_t)


class Trainer:
    def __init__(self, model, lr: float = 1e-3, dist_ctx=None, optimizer: str = "lazy"):
        """``optimizer``: "lazy" (default) = torch fused Adam on the dense weights + row-wise lazy Adam with a lazy L2
        term on the embedding arenas; "keras" = the reference's own dense Adam on everything (``KerasAdam``; single
        GPU, eager steps only)."""
        if optimizer not in ("lazy", "keras"):
            raise ValueError("optimizer must be 'lazy' or 'keras'")
        if optimizer == "keras" and dist_ctx is not None:
            raise ValueError("optimizer='keras' is the single-GPU parity mode")
        self.optimizer = optimizer
        self.keras_opt = KerasAdam(lr=lr) if optimizer == "keras" else None
        self.model = model
        self.dist = dist_ctx
        dense = model.dense_parameters()
        self._dense_params = dense
        self.dense_opt = None
        self.lr = lr
        self.sparse_opts = []
        for emb in (model.sparse_embed, model.linear_embed):
            if emb is not None:
                self.sparse_opts.append(SparseAdam(emb.arena, lr=lr, l2=emb.emb_reg))
        self.is_sigmoid = isinstance(model, XDeepFM)
        # routing sort of the embedding backward on a side stream at lookup time -- except next to the
        # persistent tcgen05 CIN kernels, which lose more to the co-scheduled sort kernels than the overlap hides
        # KON_PRESORT_X=0: keep xDeepFM's routing sort in the backward (the round-1 setting, when the sort contended
        # with the persistent CIN kernels; the forward now joins the side stream before the CIN, XDeepFM.logit)
        self.presort = not isinstance(model, XDeepFM) or os.environ.get("KON_PRESORT_X", "1") != "0"
        self._g = None
        self.capture_error = None
        # sharded jobs: a peer-barrier timeout must not pass silently (the consumer kernels would read partially
        # written rows).  Checking costs a device sync, so it runs on a cadence, after graph warm-up, and on demand.
        self.peer_check_every = 64
        self._n_steps = 0
        # measured: DeepFM 0.753 -> 0.737 ms, DCN 1.241 -> 1.213 ms; xDeepFM (a 2.3 M-parameter dense Adam next to a
        # 7 ms step) 7.16 -> 7.17 ms, i.e. nothing: left off there
        self.overlap_dense_opt = os.environ.get("KON_OVERLAP_DENSE_OPT", "0" if isinstance(model, XDeepFM) else "1") != "0"

    def check_peers(self):
        for emb in (self.model.sparse_embed, self.model.linear_embed):
            if emb is not None and hasattr(emb, "check_peer"):
                emb.check_peer()

    def _ensure_dense_opt(self):
        if self.dense_opt is None:      # lazily: layers build their weights on first call
            self._dense_params = self.model.dense_parameters()
            self.dense_opt = torch.optim.Adam(self._dense_params, lr=self.lr, eps=1e-7, fused=True,
                                              capturable=True)

    def _keras_step(self):
        """The reference's optimizer step: dense Adam on every weight, the embedding gradients densified and the
        regulariser's dense ``2 l2 w`` added (Keras adds ``l2 * sum(w^2)`` to the loss, IL:217)."""
        ko = self.keras_opt
        ko.begin_step()
        for p in self._dense_params:
            if p.grad is not None:
                ko.apply(p.data, p.grad)
        for so in self.sparse_opts:
            w = so.arena.data
            g = torch.zeros_like(w)
            for sg in getattr(so.arena, "kon_sparse_grads", None) or []:
                g += sg.to_dense(w.shape[0])
            if so.l2:
                g.add_(w, alpha=2.0 * so.l2)
            ko.apply(w, g)
            so.arena.kon_sparse_grads = []

    def loss(self, dense, ids, labels):
        out = self.model(dense, ids)
        return ops.binary_crossentropy(labels, out)        # compile(loss=binary_crossentropy), fused (kon_bce_fwd/bwd)

    def step(self, dense, ids, labels) -> torch.Tensor:
        """labels: one-hot ``[B,2]`` (``to_categorical``, DP:359) for the softmax(2) heads,
        ``[B,1]`` for XDeepFM's sigmoid head.  Returns the (device) loss."""
        for so in self.sparse_opts:
            so.arena.kon_sparse_grads = []
        ops.new_step(presort=self.presort)
        loss = self.loss(dense, ids, labels)
        self._ensure_dense_opt()
        self.dense_opt.zero_grad(set_to_none=True)
        if self.dist is not None:
            # global loss = mean over ranks of the local means; grads are summed across ranks.  The dense all-reduce
            # starts when the first embedding backward starts (all dense grads are final by then) and overlaps the
            # embedding exchange + scatter-add on a side stream.
            self.dist.arm_overlapped_allreduce(self.model, lambda: self._dense_params)
            (loss / self.dist.world).backward()
            ops.flush_deferred()
            for so in self.sparse_opts:         # the row-wise Adam does not wait for the dense all-reduce
                so.step()
            self.dist.finish_allreduce(self.model, self._dense_params)
            self.dense_opt.step()
        elif self.keras_opt is not None:
            loss.backward()
            ops.flush_deferred()
            self._keras_step()
        else:
            loss.backward()
            ops.flush_deferred()                # a first-order gradient parked for an embedding backward that never came
            # torch's fused Adam over ~10 small tensors is one 9-block launch of ~40 us (a chunk of 65,536 elements per
            # block): it runs on the side stream next to the row-wise Adam kernels of the tables instead of before them
            cur = torch.cuda.current_stream()
            side = ops._side_stream(cur.device) if (self._dense_params and self._dense_params[0].is_cuda) else None
            if side is not None and self.overlap_dense_opt:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self.dense_opt.step()
                for so in self.sparse_opts:
                    so.step()
                cur.wait_stream(side)
            else:
                self.dense_opt.step()
                for so in self.sparse_opts:
                    so.step()
        ops.end_step()
        self._n_steps += 1
        if self.dist is not None and self.peer_check_every and self._n_steps % self.peer_check_every == 0 \
                and not torch.cuda.is_current_stream_capturing():
            self.check_peers()
        return loss.detach()

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole step (launch-bound tails: ~150 kernels per step)
    # ------------------------------------------------------------------------------------------
    def capture(self, dense, ids, labels, warmup: int = 3):
        """Capture ``step`` for this batch shape; afterwards ``step_graph`` copies a batch into the
        static buffers and replays.  Returns False (and stays eager) if the capture fails."""
        self._g = None
        if self.keras_opt is not None:
            self.capture_error = "optimizer='keras' keeps its step count on the host: eager steps only"
            return False
        self._static = tuple(t.clone() for t in (dense, ids, labels))
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self.step(*self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if self.dist is not None:
                self.check_peers()
            g = torch.cuda.CUDAGraph()
            # multi-GPU: the process group's watchdog thread issues CUDA calls of its own; only THIS
            # thread's calls may invalidate the capture
            mode = "thread_local" if self.dist is not None else "global"
            with torch.cuda.graph(g, capture_error_mode=mode):
                self._static_loss = self.step(*self._static)
            self._g = g
            return True
        except Exception as e:          # noqa: BLE001 -- report and fall back to eager steps
            self._g = None
            self.capture_error = repr(e)
            torch.cuda.synchronize()
            return False

    def release_graph(self):
        """Drop the captured graph (and its private memory pool).  Call before
        ``dist.destroy_process_group()``: a live graph that holds NCCL kernels keeps the communicator busy."""
        self._g = None
        self._static = None
        self._static_loss = None

    def step_graph(self, dense, ids, labels) -> torch.Tensor:
        if self._g is None:
            return self.step(dense, ids, labels)
        for dst, src in zip(self._static, (dense, ids, labels)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._g.replay()
        self._n_steps += 1
        if self.dist is not None and self.peer_check_every and self._n_steps % self.peer_check_every == 0:
            self.check_peers()
        return self._static_loss
