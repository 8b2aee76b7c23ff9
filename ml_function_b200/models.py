"""The reference's CTR model builders on the B200 kernels: ``FM``, ``DeepFM``, ``DCN``,
``XDeepFM``, ``AutoInt`` (kon/model/ctr_model/model/models.py = MD:36-41, 80-106, 121-138,
150-165) and ``data_prepare.FeatureInput`` / ``InputFeature`` (kon/utils/data_prepare.py =
DP:39-76).

A builder takes an ``InputFeature`` exactly as in the reference and returns a module whose
``forward(dense_inputs, sparse_inputs)`` accepts either the reference's lists of per-feature
``[B,1]`` tensors or packed ``dense [B,n_dense]`` float32 / ``ids [B,F]`` int32|int64.

Data layout in HBM: one row-major concat buffer ``xcat [B, W]`` per step,
``W = F*k + n_dense`` rounded up to a multiple of 4 floats.  Columns ``[0, F*k)`` hold the
field embeddings (written in place by the gather kernel, 16-B aligned rows), then the dense
features, then zero padding.  ``xcat[:, :F*k].view(B,F,k)`` is the FM / CIN / attention
input and ``xcat`` itself the MLP / Cross input -- no concat copy.  The reference's
``StackLayer`` order is dense-first (MD:86); weights in reference order are permuted once at
load time (``load_reference_params``), which changes nothing but the physical column order.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from .layers import (CIN, AttentionBaseLayer, CrossLayer, DnnLayer, FieldList, FmLayer, InnerLayer, IPnnLayer, MergeScoreLayer,
                     MultHeadAttentionLayer, ProductAttentionLayer, ScoreLayer, SeqBaseLayer, SparseEmbed, StackLayer,
                     denseFea, pack_ids, sparseFea)


class InputFeature(object):
    """DP:39-49 (fields kept; the Keras Input placeholders have no torch counterpart, the
    embed entries are the ``SparseEmbed`` layers themselves)."""

    def __init__(self, denseInfo=None, sparseInfo=None, seqInfo=None, denseInputs=None, sparseInputs=None,
                 seqInputs=None, linearEmbed=None, sparseEmbed=None, seqEmbedList=None):
        self.dense_info = denseInfo
        self.sparse_info = sparseInfo
        self.seq_info = seqInfo
        self.dense_inputs = denseInputs
        self.sparse_inputs = sparseInputs
        self.seq_inputs = seqInputs
        self.linear_embed = linearEmbed
        self.sparse_embed = sparseEmbed
        self.seq_embed_list = seqEmbedList


def FeatureInput(sparseInfo: list = None, denseInfo: list = None, seqInfo=None, useLinear: bool = False,
                 useAddLinear: bool = False, useFlattenLinear: bool = False, useFlattenSparse: bool = False,
                 device="cuda") -> InputFeature:
    """``data_prepare.FeatureInput`` (DP:65-76)."""
    linearEmbed = sparseEmbed = None
    if useLinear:
        linearEmbed = SparseEmbed(sparseInfo, use_flatten=useFlattenLinear, is_linear=True,
                                  use_add=useAddLinear, device=device)
    if sparseInfo:
        sparseEmbed = SparseEmbed(sparseInfo, use_flatten=useFlattenSparse, device=device)
    return InputFeature(denseInfo, sparseInfo, seqInfo, None, None, None, linearEmbed, sparseEmbed, [None, None])


def _pack_dense(dense_inputs) -> torch.Tensor:
    if isinstance(dense_inputs, torch.Tensor):
        return dense_inputs
    return torch.cat([t.reshape(t.shape[0], 1) for t in dense_inputs], dim=1)


class _CtrModel(nn.Module):
    """Shared front end: ids/dense -> the concat buffer and its views."""

    def __init__(self, inputFea: InputFeature):
        super().__init__()
        self.sparse_embed: SparseEmbed = inputFea.sparse_embed
        self.linear_embed: Optional[SparseEmbed] = inputFea.linear_embed
        self.n_dense = len(inputFea.dense_info or [])
        self.F = len(inputFea.sparse_info)
        self.k = self.sparse_embed.dim
        self.Fk = self.F * self.k
        self.W = (self.Fk + self.n_dense + 3) // 4 * 4

    # ---- physical <-> reference column order -------------------------------------------
    def ref_to_phys_rows(self, w_ref: torch.Tensor) -> torch.Tensor:
        """Rows of a reference-order ``[n_dense + F*k, ...]`` weight -> physical ``[W, ...]``."""
        nd = self.n_dense
        pad = torch.zeros((self.W - self.Fk - nd,) + tuple(w_ref.shape[1:]), dtype=w_ref.dtype, device=w_ref.device)
        return torch.cat([w_ref[nd:], w_ref[:nd], pad], dim=0).contiguous()

    def phys_to_ref_rows(self, w_phys: torch.Tensor) -> torch.Tensor:
        """Inverse of ``ref_to_phys_rows`` (drops the pad rows)."""
        nd = self.n_dense
        return torch.cat([w_phys[self.Fk:self.Fk + nd], w_phys[:self.Fk]], dim=0)

    # ---- gradients under the reference's weight names / layouts (parity tests, checkpoints) ----
    def _dnn_grads(self, out: Dict[str, torch.Tensor], dnn: "DnnLayer", first_is_xcat: bool):
        for i, (w, b) in enumerate(zip(dnn.kernels, dnn.biases)):
            if w.grad is not None:
                out[f"dnn_w{i}"] = self.phys_to_ref_rows(w.grad) if (i == 0 and first_is_xcat) else w.grad
            if b.grad is not None:
                out[f"dnn_b{i}"] = b.grad
        if dnn.logit_kernel is not None and dnn.logit_kernel.grad is not None:
            out["dnn_logit_w"], out["dnn_logit_b"] = dnn.logit_kernel.grad, dnn.logit_bias.grad

    def reference_grads(self) -> Dict[str, torch.Tensor]:
        """Dense-parameter gradients keyed by the names ``load_reference_params`` takes, in the reference's
        layouts (rows of first-layer kernels back in dense-first order, ``[D,1]`` cross kernels, ...)."""
        raise NotImplementedError

    def front(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        dense = _pack_dense(dense_inputs) if self.n_dense else None
        xcat = self.sparse_embed.lookup_concat(ids, dense, self.W)
        v = xcat[:, :self.Fk].view(ids.shape[0], self.F, self.k)
        return ids, xcat, v

    def sparse_parameters(self) -> List[nn.Parameter]:
        ps = [self.sparse_embed.arena]
        if self.linear_embed is not None:
            ps.append(self.linear_embed.arena)
        return ps

    def dense_parameters(self) -> List[nn.Parameter]:
        sp = {id(p) for p in self.sparse_parameters()}
        return [p for p in self.parameters() if id(p) not in sp]


class FM(_CtrModel):
    """MD:36-41: ``Dense(2,softmax)(squeeze(FmLayer([sparse_embed, linear_embed])))``."""

    def __init__(self, inputFea: InputFeature = None):
        super().__init__(inputFea)
        self.fm = FmLayer()
        self.head = MergeScoreLayer(use_merge=False)

    def forward(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        v = self.sparse_embed.lookup(ids)
        lin = self.linear_embed.lookup_sum(ids)            # FmLayer only uses sum_f lin[b,f]
        fm_ = self.fm([v, lin])
        return self.head(fm_.squeeze(1))

    def load_reference_params(self, p: Dict[str, torch.Tensor]):
        _load_embeds(self, p)
        self.head.load_reference_weights(p["head_w"], p["head_b"])

    def reference_grads(self):
        return {"head_w": self.head.kernel.grad, "head_b": self.head.bias.grad}


class DeepFM(_CtrModel):
    """MD:80-90."""

    def __init__(self, inputFea: InputFeature = None, hidden_units=None):
        super().__init__(inputFea)
        self.hidden_units = hidden_units if hidden_units is not None else [256, 128, 64]
        self.fm = FmLayer()
        self.dnn = DnnLayer(hidden_units=self.hidden_units)
        self.head = MergeScoreLayer()

    def forward(self, dense_inputs, sparse_inputs):
        ids, xcat, v = self.front(dense_inputs, sparse_inputs)
        ops.two_branch_input(xcat)            # FM + MLP: the FM gradient is added into the MLP's in place (ops._XGRAD)
        lin = self.linear_embed.lookup_sum(ids)
        fm_ = self.fm.on_concat(xcat, lin, self.F, self.k)     # = self.fm([v, lin]) on the window of xcat
        dnn_ = self.dnn(xcat)
        return self.head([fm_, dnn_])

    def load_reference_params(self, p):
        _load_embeds(self, p)
        n = len(self.hidden_units)
        ks = [p[f"dnn_w{i}"] for i in range(n)]
        ks[0] = self.ref_to_phys_rows(ks[0])
        self.dnn.load_reference_weights(ks, [p[f"dnn_b{i}"] for i in range(n)])
        self.head.load_reference_weights(p["head_w"], p["head_b"])

    def reference_grads(self):
        out = {"head_w": self.head.kernel.grad, "head_b": self.head.bias.grad}
        self._dnn_grads(out, self.dnn, True)
        return out


class DCN(_CtrModel):
    """MD:92-106."""

    def __init__(self, inputFea: InputFeature = None, hidden_units=None, cross_hidden=3):
        super().__init__(inputFea)
        self.hidden_units = hidden_units if hidden_units is not None else [256, 128, 64]
        self.cross = CrossLayer(cross_hidden=cross_hidden, n_valid=self.Fk + self.n_dense)
        self.dnn = DnnLayer(hidden_units=self.hidden_units)
        self.head = MergeScoreLayer()

    def forward(self, dense_inputs, sparse_inputs):
        ids, xcat, v = self.front(dense_inputs, sparse_inputs)
        ops.two_branch_input(xcat)            # cross + MLP
        cross_fea = self.cross(xcat)                       # [B,W,1]
        deep_fea = self.dnn(xcat)
        return self.head([cross_fea, deep_fea])

    def load_reference_params(self, p):
        _load_embeds(self, p)
        n, L = len(self.hidden_units), self.cross.cross_hidden
        ks = [p[f"dnn_w{i}"] for i in range(n)]
        ks[0] = self.ref_to_phys_rows(ks[0])
        self.dnn.load_reference_weights(ks, [p[f"dnn_b{i}"] for i in range(n)])
        self.cross.load_reference_weights([self.ref_to_phys_rows(p[f"outer_weight_{i}"]) for i in range(L)],
                                          [self.ref_to_phys_rows(p[f"outer_bias_{i}"]) for i in range(L)])
        D = self.n_dense + self.Fk
        hw = p["head_w"]
        self.head.load_reference_weights(torch.cat([self.ref_to_phys_rows(hw[:D]), hw[D:]], 0), p["head_b"])

    def reference_grads(self):
        hg = self.head.kernel.grad
        out = {"head_w": torch.cat([self.phys_to_ref_rows(hg[:self.W]), hg[self.W:]], 0), "head_b": self.head.bias.grad}
        for i in range(self.cross.cross_hidden):
            out[f"outer_weight_{i}"] = self.phys_to_ref_rows(self.cross.kernel.grad[i].unsqueeze(1))
            out[f"outer_bias_{i}"] = self.phys_to_ref_rows(self.cross.bias.grad[i].unsqueeze(1))
        self._dnn_grads(out, self.dnn, True)
        return out


class XDeepFM(_CtrModel):
    """MD:121-138, with ``FeatureInput(useLinear=True, useAddLinear=True)`` (the only wiring
    under which ``ScoreLayer(use_add=True)([linear, cin, dnn])`` is well formed).  Output
    ``sigmoid`` ``[B,1,1]``."""

    def __init__(self, inputFea: InputFeature = None, conv_size=None, hidden_units=None, cin_precision="bf16"):
        super().__init__(inputFea)
        self.conv_size = conv_size if conv_size is not None else [200, 200, 200]
        self.hidden_units = hidden_units if hidden_units is not None else [256, 128, 64]
        self.cin = CIN(conv_size=self.conv_size, output_dim=1, precision=cin_precision)
        self.dnn = DnnLayer(hidden_units=self.hidden_units, output_dim=1)
        self.score = ScoreLayer(use_add=True)

    def logit(self, dense_inputs, sparse_inputs):
        ids, xcat, v = self.front(dense_inputs, sparse_inputs)
        linear = self.linear_embed.lookup_sum(ids)         # [B,1] (useAddLinear, IL:233-234)
        # the MLP runs BEFORE the CIN (the sum below keeps the reference's operand order): its short kernels share
        # the GPU with the backward's routing sort on the side stream; the persistent CIN kernels then start on an
        # idle GPU (ops.join_presort)
        dnn_out = self.dnn(xcat)                           # [B,1]
        ops.join_presort()
        cin_out = self.cin(xcat, fields=(self.F, self.k))  # [B,1]; reads xcat[:, :F*k] in place
        return ScoreLayer.summed([linear.unsqueeze(1), cin_out, dnn_out])   # [B,1,1]

    def forward(self, dense_inputs, sparse_inputs):
        return torch.sigmoid(self.logit(dense_inputs, sparse_inputs))

    def load_reference_params(self, p):
        _load_embeds(self, p)
        n, nc = len(self.hidden_units), len(self.conv_size)
        ks = [p[f"dnn_w{i}"] for i in range(n)]
        ks[0] = self.ref_to_phys_rows(ks[0])
        self.dnn.load_reference_weights(ks, [p[f"dnn_b{i}"] for i in range(n)], p["dnn_logit_w"], p["dnn_logit_b"])
        self.cin.load_reference_weights([p[f"cin_w{i}"] for i in range(nc)], [p[f"cin_b{i}"] for i in range(nc)],
                                        p["cin_logit_w"], p["cin_logit_b"])

    def reference_grads(self):
        out = {"cin_logit_w": self.cin.logit_kernel.grad, "cin_logit_b": self.cin.logit_bias.grad}
        for i, (w, b) in enumerate(zip(self.cin.conv_kernels, self.cin.conv_biases)):
            out[f"cin_w{i}"], out[f"cin_b{i}"] = w.grad, b.grad
        self._dnn_grads(out, self.dnn, True)
        return out


class NFM(_CtrModel):
    """MD:108-119 (a sibling that reuses the FM kernel).  The bi-interaction
    ``InnerLayer(use_inner=True, use_add=True)`` is ``kon_fm_fwd/bwd`` without the linear term; the MLP
    sees ``[dense | bi-interaction]`` (reference order, MD:113); output ``sigmoid`` ``[B,1,1]``."""

    def __init__(self, inputFea: InputFeature = None, hidden_units=None):
        super().__init__(inputFea)
        self.hidden_units = hidden_units if hidden_units is not None else [256, 128, 64]
        self.inner = InnerLayer(use_inner=True, use_add=True)
        self.dnn = DnnLayer(hidden_units=self.hidden_units, output_dim=1)

    def logit(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        v = self.sparse_embed.lookup(ids)                  # [B,F,k]
        cross = self.inner(v)                              # [B,1,k]
        linear = self.linear_embed.lookup_sum(ids)         # [B,1] = Add over the F first-order terms
        parts = [cross.reshape(cross.shape[0], -1)]
        if self.n_dense:
            parts.insert(0, _pack_dense(dense_inputs))
        dnn_out = self.dnn(torch.cat(parts, dim=1))        # [B,1]
        return ScoreLayer.summed([linear.unsqueeze(1), dnn_out])      # [B,1,1]

    def forward(self, dense_inputs, sparse_inputs):
        return torch.sigmoid(self.logit(dense_inputs, sparse_inputs))

    def load_reference_params(self, p):
        _load_embeds(self, p)
        n = len(self.hidden_units)
        self.dnn.load_reference_weights([p[f"dnn_w{i}"] for i in range(n)], [p[f"dnn_b{i}"] for i in range(n)],
                                        p["dnn_logit_w"], p["dnn_logit_b"])

    def reference_grads(self):
        out = {}
        self._dnn_grads(out, self.dnn, False)
        return out


class AFM(_CtrModel):
    """MD:141-147: ``InnerLayer()`` pairwise products -> ``AttentionBaseLayer()`` ->
    ``ScoreLayer(use_add=True)(linear_embed + [atten_output])`` = sigmoid ``[B,1,1]``.  The products are
    never materialised here: the pooling weights of the reference are identically 1 (see
    ``AttentionBaseLayer``), so the pooled vector is the FM kernel's ``sum_{i<j} v_i v_j``."""

    def __init__(self, inputFea: InputFeature = None):
        super().__init__(inputFea)
        self.inner = InnerLayer()
        self.atten = AttentionBaseLayer()

    def logit(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        v = self.sparse_embed.lookup(ids)                  # [B,F,k]
        atten_out = self.atten.fused(v)                    # [B,1]
        linear = self.linear_embed.lookup_sum(ids)         # Add over the F first-order [B,1,1] terms
        return ScoreLayer.summed([linear.unsqueeze(1), atten_out])      # [B,1,1]

    def forward(self, dense_inputs, sparse_inputs):
        return torch.sigmoid(self.logit(dense_inputs, sparse_inputs))

    def load_reference_params(self, p):
        _load_embeds(self, p)
        self.atten.load_reference_weights(p["afm_score_w"], p["afm_score_b"], p["afm_mlp_w"], p["afm_out_w"], p["afm_out_b"])

    def reference_grads(self):
        a = self.atten
        z = lambda w: torch.zeros_like(w) if w.grad is None else w.grad     # never-read scoring weights
        return {"afm_score_w": z(a.kernel_w), "afm_score_b": z(a.kernel_b), "afm_mlp_w": z(a.mlp_kernel),
                "afm_out_w": a.out_kernel.grad, "afm_out_b": a.out_bias.grad}


class PNN(_CtrModel):
    """MD:43-56 with ``use_inner=True, use_outer=False`` (IPNN; ``OPnnLayer`` is broken in the reference itself, IL:56 vs
    IL:63, and is rejected): ``StackLayer(linear_embed + IPnnLayer()(sparse_embed))`` -> ``DnnLayer`` ->
    ``Dense(2, softmax)``.  The pairwise products come from ``kon_pairs_fwd/bwd``."""

    def __init__(self, inputFea: InputFeature = None, hidden_units=None, use_inner=True, use_outer=False):
        super().__init__(inputFea)
        if use_outer or not use_inner:
            raise ValueError("PNN: only use_inner=True, use_outer=False can run (OPnnLayer is broken in the reference)")
        self.hidden_units = hidden_units if hidden_units is not None else [256, 256, 256]
        self.ipnn = IPnnLayer()
        self.dnn = DnnLayer(hidden_units=self.hidden_units)
        self.head = MergeScoreLayer(use_merge=False)

    def forward(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        v = self.sparse_embed.lookup(ids)                  # [B,F,k]
        lin = self.linear_embed.lookup(ids)                # [B,F,1]: the 26 first-order terms, un-summed (MD:48)
        pairs = self.ipnn(v).packed                        # [B,P,k]
        x = torch.cat([lin.reshape(lin.shape[0], -1), pairs.reshape(pairs.shape[0], -1)], dim=1)   # StackLayer (MD:53)
        return self.head(self.dnn(x))

    def load_reference_params(self, p):
        _load_embeds(self, p)
        n = len(self.hidden_units)
        self.dnn.load_reference_weights([p[f"dnn_w{i}"] for i in range(n)], [p[f"dnn_b{i}"] for i in range(n)])
        self.head.load_reference_weights(p["head_w"], p["head_b"])

    def reference_grads(self):
        out = {"head_w": self.head.kernel.grad, "head_b": self.head.bias.grad}
        self._dnn_grads(out, self.dnn, False)
        return out


class AutoInt(_CtrModel):
    """MD:150-165.  ``n_layers > 1`` stacks blocks by re-packing ``[H,B,F,d] -> [B,F,H*d]``
    (standard AutoInt; an extension, the reference wires exactly one block)."""

    def __init__(self, inputFea: InputFeature = None, attention_dim=8, attention_head_dim=3, n_layers=1,
                 precision="fp32"):
        super().__init__(inputFea)
        self.blocks = nn.ModuleList([
            DnnLayer(res_unit=1, other_dense=[MultHeadAttentionLayer(
                attention_dim=attention_dim, attention_head_dim=attention_head_dim, use_ln=True,
                atten_mask_mod=1, precision=precision)]) for _ in range(n_layers)])
        self.head = MergeScoreLayer(use_merge=False)

    def forward(self, dense_inputs, sparse_inputs):
        ids = pack_ids(sparse_inputs)
        x = self.sparse_embed.lookup(ids)                  # StackLayer(use_flat=False, axis=1)
        for i, blk in enumerate(self.blocks):
            last = i + 1 == len(self.blocks)
            # [H,B,F,d]; on the bf16 path the kernel writes it in the order its consumer reads, so the
            # re-packing below is a view (no permute copy, forward or backward)
            a = blk(x, layout="bhfd" if last else "bfhd")
            if not last:
                x = a.permute(1, 2, 0, 3).reshape(a.shape[1], a.shape[2], -1)
                x = x if x.is_contiguous() else x.contiguous()
        final = a.permute(1, 0, 2, 3).reshape(a.shape[1], -1)   # MD:162: heads side by side
        return self.head(final)

    def load_reference_params(self, p):
        """Block 0 takes the reference's names (``query_w`` ...); stacked blocks l >= 1 (the extension) take the
        same names with an ``_l`` suffix (``query_w_1`` ...)."""
        _load_embeds(self, p)
        for l, blk in enumerate(self.blocks):
            sfx = "" if l == 0 else f"_{l}"
            if ("query_w" + sfx) not in p:
                if l == 0:
                    raise KeyError("query_w")
                continue
            blk.other_dense[0].load_reference_weights(p["query_w" + sfx], p["key_w" + sfx], p["res_w" + sfx],
                                                      p["ln_gamma" + sfx], p["ln_beta" + sfx],
                                                      value_w=p.get("value_w" + sfx))
        self.head.load_reference_weights(p["head_w"], p["head_b"])

    def reference_grads(self):
        out = {"head_w": self.head.kernel.grad, "head_b": self.head.bias.grad}
        for l, blk in enumerate(self.blocks):
            sfx = "" if l == 0 else f"_{l}"
            att = blk.other_dense[0]
            for n in ("query_w", "key_w", "res_w", "ln_gamma", "ln_beta"):
                out[n + sfx] = getattr(att, n).grad
        return out


def _load_embeds(model: _CtrModel, p):
    F = model.F
    model.sparse_embed.load_reference_weights([p[f"emb_{f}"] for f in range(F)])
    if model.linear_embed is not None:
        model.linear_embed.load_reference_weights([p[f"lin_{f}"] for f in range(F)])


def keras_binary_crossentropy(y_true: torch.Tensor, y_pred: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """``compile(loss=binary_crossentropy)`` (example/ctr_example/un_seq.py:61) on
    probabilities: clip, mean over the last axis, mean over the batch."""
    p = torch.clamp(y_pred, eps, 1 - eps)
    bce = -(y_true * torch.log(p + eps) + (1 - y_true) * torch.log(1 - p + eps))
    return bce.mean(dim=-1).mean()
