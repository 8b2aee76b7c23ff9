"""Host-side mirror of the reference's Keras layers for the CTR hot path.

Same class names, constructor keywords, input list structure and output shapes as
``kon/model/ctr_model/layer/{interactive_layer,behavior_layer,core_layer}`` (cited per
class as IL / BL / CL), on ``torch.Tensor`` (CUDA) instead of ``tf.Tensor``.  Every
hot-path op is a call into ``libkon_b200.so`` (``ops.py``); torch provides parameters,
autograd plumbing and the cuBLAS ``Dense`` layers of the adjacent MLP / heads (SURVEY
§8 a12: not a custom-kernel target).  Nothing here computes on the CPU: CPU tensors are
rejected by the library (``KonError``).

Weights keep the reference names/shapes (``outer_weight_i [D,1]``, ``query_w [k,H,d]``,
Conv1D kernel ``[1,C,N]``, Embedding ``[word_size,dim]``) so reference-layout weights
load verbatim through the ``load_reference_weights`` helpers.
"""
from __future__ import annotations

from collections import namedtuple
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import ops

# DP:59-60 (data_prepare.__init__): the feature descriptors the layers are built from.
sparseFea = namedtuple('sparseFea', ['fea_name', 'word_size', 'input_dim', 'cross_unit', 'linear_unit',
                                     'pre_weight', 'mask_zero', 'is_trainable', 'input_length',
                                     'sample_num', 'batch_size', 'emb_reg'])
denseFea = namedtuple('denseFea', ['fea_name', 'batch_size'])


def make_sparse_fea(name, word_size, cross_unit=8, linear_unit=1, input_length=1, mask_zero=False,
                    emb_reg=1e-8):
    """``sparse_fea_deal`` defaults (DP:85-102)."""
    return sparseFea(name, word_size, word_size, cross_unit, linear_unit, None, mask_zero, True,
                     input_length, None, None, emb_reg)


class FieldList(list):
    """The list of per-field tensors the reference passes between layers, carrying the packed
    ``[B,F,k]`` buffer the views were cut from (``.packed``) so that the next layer can hand
    the whole block to one kernel without a concat copy."""

    packed: Optional[torch.Tensor] = None

    @staticmethod
    def of(packed: torch.Tensor, flatten: bool = False) -> "FieldList":
        F = packed.shape[1]
        fl = FieldList(packed[:, f] if flatten else packed[:, f:f + 1] for f in range(F))
        fl.packed = packed
        return fl


def pack_fields(x) -> torch.Tensor:
    """list of ``[B,1,k]`` / ``[B,k]`` (or an already packed ``[B,F,k]``) -> ``[B,F,k]``."""
    if isinstance(x, torch.Tensor):
        return x
    p = getattr(x, "packed", None)
    if p is not None:
        return p
    xs = [t if t.dim() == 3 else t.unsqueeze(1) for t in x]
    return torch.cat(xs, dim=1)


def pack_ids(inputs) -> torch.Tensor:
    """Reference call convention: a list of ``[B, input_length]`` id tensors (one Keras Input
    per field, DP:317-333).  Packed ``[B,F]`` / ``[B,F,L]`` integer tensors pass through.
    The reference feeds ids as float32 (DP:290-292) and Keras casts to int32; float ids are
    cast the same way here."""
    if isinstance(inputs, torch.Tensor):
        ids = inputs
    else:
        cols = [t if t.dim() == 2 else t.unsqueeze(1) for t in inputs]
        L_ = cols[0].shape[1]
        ids = torch.stack(cols, dim=1)            # [B,F,L]
        if L_ == 1:
            ids = ids[:, :, 0]
    if ids.dtype not in (torch.int32, torch.int64):
        ids = ids.to(torch.int32)
    return ids.contiguous()


# --------------------------------------------------------------------------------------
# a1-a4  SparseEmbed (IL:189-247), SeqBaseLayer (BL:32-51)
# --------------------------------------------------------------------------------------
class SparseEmbed(nn.Module):
    """IL:196.  All fields' tables live back to back in ONE ``arena [R, dim]`` parameter
    (field f's rows start at ``field_row_offset[f]``), so the 26 Keras ``Embedding`` gathers
    are one kernel.  ``call`` returns the reference's list of ``[B, input_length, dim]``
    tensors (``[B, input_length*dim]`` when ``use_flatten``; one tensor when ``use_add``),
    as views of one packed buffer.

    The gradient of ``arena`` is sparse (``arena.kon_sparse_grads``); ``arena.grad`` stays
    ``None``.  ``emb_reg`` (IL:217) is applied lazily by the sparse optimizers."""

    def __init__(self, sparse_info: list, is_linear=False, use_flatten=True, use_add=False, seed=2020,
                 support_masking=True, mask_zero=False, device="cuda"):
        super().__init__()
        self.sparse_info = list(sparse_info)
        self.is_linear = is_linear
        self.use_flatten = use_flatten
        self.use_add = use_add
        self.seed = seed
        self.supports_masking = support_masking
        self.mask_zero = mask_zero
        dims = {(i.linear_unit if is_linear else i.cross_unit) for i in self.sparse_info}
        if len(dims) != 1:
            raise ValueError("SparseEmbed: all fields must share one embedding dim, got %s" % sorted(dims))
        self.dim = dims.pop()
        rows = [int(i.word_size) for i in self.sparse_info]
        offs = [0]
        for r in rows:
            offs.append(offs[-1] + r)
        self.field_row_offset = tuple(offs)
        self.emb_reg = 0.0 if is_linear else float(self.sparse_info[0].emb_reg or 0.0)
        arena = torch.empty(offs[-1], self.dim, device=device, dtype=torch.float32)
        if arena.device.type != "meta":       # "meta": a placeholder that parallel.DistContext.attach replaces by shards
            # per-table init, chunked on the device: glorot_uniform for cross tables (IL:215),
            # Keras Embedding default 'uniform' = U(-0.05,0.05) for the linear ones (IL:220-222)
            gd = torch.Generator(device=device).manual_seed(seed)
            for f, r in enumerate(rows):
                lim = 0.05 if is_linear else (6.0 / (r + self.dim)) ** 0.5
                arena[offs[f]:offs[f + 1]].uniform_(-lim, lim, generator=gd)
        self.arena = nn.Parameter(arena)

    def load_reference_weights(self, tables: Sequence[torch.Tensor]):
        """``tables[f]``: the Keras Embedding matrix ``[word_size_f, dim]`` of field f."""
        with torch.no_grad():
            for f, t in enumerate(tables):
                lo, hi = self.field_row_offset[f], self.field_row_offset[f + 1]
                assert tuple(t.shape) == (hi - lo, self.dim), (t.shape, hi - lo, self.dim)
                self.arena[lo:hi].copy_(t)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        """packed ids ``[B,F]`` / ``[B,F,L]`` -> ``[B,F,dim]`` (bag-summed over L)."""
        return ops.embed_lookup(self.arena, ids, self.field_row_offset, False)

    def lookup_sum(self, ids: torch.Tensor) -> torch.Tensor:
        """packed ids -> ``[B,dim]`` = sum over fields, left to right (Keras Add, IL:233-234)."""
        return ops.embed_lookup(self.arena, ids, self.field_row_offset, True)

    def lookup_concat(self, ids: torch.Tensor, dense: Optional[torch.Tensor], width: int) -> torch.Tensor:
        """-> ``xcat [B,width]`` = field embeddings | dense features | zero pad, the gather
        kernel writing the embedding window in place (CL:49-55 without the concat copy)."""
        return ops.embed_lookup_concat(self.arena, ids, self.field_row_offset, dense, width)

    def forward(self, inputs, **kwargs):
        ids = pack_ids(inputs)
        if self.use_add:                                   # IL:233-234 (input_length 1)
            if ids.dim() == 3 and ids.shape[2] > 1:
                raise L.KonError("SparseEmbed(use_add=True) with input_length > 1 is not provided")
            out = ops.embed_lookup(self.arena, ids, self.field_row_offset, True)   # [B,dim]
            return out if self.use_flatten else out.unsqueeze(1)
        if ids.dim() == 3 and ids.shape[2] > 1:
            # sequence features (DP:74): the reference returns the un-pooled [B,L,dim] per field (and the
            # Embedding masks) and pools later in SeqBaseLayer.  Same gather kernel, one row per (b,l,f):
            # the ids are re-laid as [B*L, F] so that the output is one [B,L,F,dim] buffer whose per-field
            # slices are the reference's tensors.  (Gather + pool in ONE pass: SeqBaseLayer.fused / lookup().)
            B, F, L_ = ids.shape
            ids_blf = ids.permute(0, 2, 1).reshape(B * L_, F).contiguous()
            seq = ops.embed_lookup(self.arena, ids_blf, self.field_row_offset, False).view(B, L_, F, self.dim)
            fl = FieldList(seq[:, :, f].reshape(B, L_ * self.dim) if self.use_flatten else seq[:, :, f]
                           for f in range(F))
            fl.packed_seq = seq
            if self.mask_zero:                             # IL:238-242: Embedding.compute_mask = (ids != 0)
                return fl, [ids[:, f] != 0 for f in range(F)]
            return fl
        packed = self.lookup(ids)
        fl = FieldList.of(packed, flatten=self.use_flatten)
        if self.mask_zero:                                 # IL:238-242: (embeds, masks)
            masks = [ids[:, f:f + 1] != 0 for f in range(ids.shape[1])]
            return fl, masks
        return fl


class SeqBaseLayer(nn.Module):
    """BL:37-51: sum-pool ``[B,L,k]`` sequence embeddings over L.  ``fused(embed, ids)`` is
    the hot path (gather + pool in one kernel); ``forward`` accepts materialised lists for
    signature compatibility and pools them with the same kernel-side order l = 0..L-1."""

    def __init__(self, supports_masking=True, mask_zero=False, **kwargs):
        super().__init__()
        self.supports_masking = supports_masking
        self.mask_zero = mask_zero

    @staticmethod
    def fused(embed: SparseEmbed, ids: torch.Tensor) -> FieldList:
        return FieldList.of(embed.lookup(pack_ids(ids)))

    def forward(self, inputs, mask=None):
        """BL:45-46 on materialised tensors: ``[expand_dims(reduce_sum(x, 1), 1) for x in inputs]``, summed in
        order l = 0..L-1 by ``kon_pool_sum_fwd``.  A list that came out of ``SparseEmbed`` is pooled in one
        launch over its packed ``[B,L,F,k]`` buffer.  Padding id 0 still contributes row 0, as in the
        reference (the mask is only forwarded, BL:48-51)."""
        seq = getattr(inputs, "packed_seq", None)
        if seq is not None:
            B, L_, F, k = seq.shape
            pooled = ops.pool_sum(seq.view(B, L_, F * k)).view(B, F, k)
            return FieldList.of(pooled)
        if isinstance(inputs, torch.Tensor):
            inputs = [inputs]
        return [ops.pool_sum(x).unsqueeze(1) for x in inputs]

    def compute_mask(self, pre_mask, mask=None):
        return pre_mask if self.mask_zero else None


# --------------------------------------------------------------------------------------
# a5-a6  InnerLayer (IL:34-66), FmLayer (IL:145-170)
# --------------------------------------------------------------------------------------
class InnerLayer(nn.Module):
    """IL:38.  ``use_inner=True, use_add=True`` (the NFM / FM use) is the fused kernel
    ``sum_{i<j} v_i * v_j``.  The un-summed list of 325 products and ``use_inner=False``
    (broken in the reference: ``self.dot`` is commented out, IL:56 vs IL:63) are outside the
    hot path."""

    def __init__(self, use_inner: bool = True, mod=1, seed=2020, perm=None, use_add=False):
        super().__init__()
        self.use_inner, self.mod, self.seed, self.perm, self.use_add = use_inner, mod, seed, perm, use_add

    def forward(self, inputs, **kwargs):
        if not self.use_inner:
            raise L.KonError("InnerLayer(use_inner=False) is broken in the reference itself (self.dot is "
                             "commented out at IL:56 but used at IL:63) and is not provided")
        v = pack_fields(inputs)
        if self.use_add:
            return ops.fm(v, None).unsqueeze(1)            # [B,1,k] = Add(pairwise products)
        # IL:61: the list of F(F-1)/2 products [B,1,k] (AFM / IPNN input), views of one packed [B,P,k] buffer
        P = ops.pairs(v)
        fl = FieldList.of(P)
        fl.source_fields = v
        return fl


class IPnnLayer(nn.Module):
    """IL:68-80: the inner-product half of PNN = ``InnerLayer()`` without the Add: the list of pairwise products."""

    def __init__(self, seed=2020):
        super().__init__()
        self.seed = seed
        self.inner = InnerLayer()

    def forward(self, inputs, **kwargs):
        return self.inner(inputs)


class FmLayer(nn.Module):
    """IL:146.  ``inputs = [cross_embed, linear_embed]`` -> ``[B,1,k]``
    (= ``Add([Add(pairwise products)] + linear_list)`` with ``[B,1,1]`` broadcast)."""

    def __init__(self, use_inner: bool = True, mod=1, use_add=True, **kwargs):
        super().__init__()
        self.cross = InnerLayer(use_inner=use_inner, mod=mod, use_add=use_add)
        self.use_add = use_add

    def forward(self, inputs, **kwargs):
        if not self.use_add:
            raise L.KonError("FmLayer(use_add=False) is not on the B200 hot path")
        v = pack_fields(inputs[0])
        lin = pack_fields(inputs[1])                       # [B,F,1] (or an already reduced [B,F'])
        lin = lin.reshape(lin.shape[0], -1)                # only the sum over fields enters the result
        return ops.fm(v, lin).unsqueeze(1)

    def on_concat(self, xcat, lin, F: int, k: int):
        """Same result for ``cross_embed = xcat[:, :F*k]`` (the model's concat buffer, a11), with the
        gradient of ``xcat`` produced in one piece."""
        lin = None if lin is None else lin.reshape(lin.shape[0], -1)
        return ops.fm_xcat(xcat, lin, F, k).unsqueeze(1)


# --------------------------------------------------------------------------------------
# a7  CrossLayer (IL:250-282)
# --------------------------------------------------------------------------------------
class CrossLayer(nn.Module):
    """IL:255.  ``[B,D] -> [B,D,1]``; weights ``outer_weight_i`` / ``outer_bias_i`` ``[D,1]``
    stored stacked as ``kernel [L,D]`` / ``bias [L,D]``."""

    def __init__(self, cross_hidden=3, seed=2020, n_valid=None, **kwargs):
        super().__init__()
        self.cross_hidden, self.seed = cross_hidden, seed
        # ``n_valid``: the reference's D when the input carries zero alignment columns behind it (the models'
        # concat buffer is padded to a multiple of 4 floats).  Pad columns are inert: x0 = 0 there, the kernels'
        # pad entries are 0 and get a zero gradient, and the pad BIAS gradient (the only thing that could make
        # them a learnt constant feature) is masked.
        self.n_valid = n_valid
        self.kernel = nn.UninitializedParameter()
        self.bias = nn.UninitializedParameter()

    def _mask_pad(self):
        D = self.bias.shape[1]
        nv = D if self.n_valid is None else self.n_valid
        if nv < D:
            mask = torch.ones(1, D, device=self.bias.device)
            mask[:, nv:] = 0
            self.bias.register_hook(lambda g: g * mask)

    def build(self, D: int, device):
        nv = D if self.n_valid is None else self.n_valid
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        lim = (6.0 / (nv + 1)) ** 0.5                      # glorot_uniform on the reference's [D,1]
        # the reference passes the same seeded initializer object to every layer (IL:267)
        w = torch.zeros(self.cross_hidden, D)
        for i in range(self.cross_hidden):
            w[i, :nv] = (torch.rand(nv, generator=g) * 2 - 1) * lim
        self.kernel = nn.Parameter(w.to(device))
        self.bias = nn.Parameter(torch.zeros(self.cross_hidden, D, device=device))
        self._mask_pad()

    def load_reference_weights(self, kernels: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
        w = torch.stack([k.reshape(-1) for k in kernels])
        b = torch.stack([k.reshape(-1) for k in biases])
        self.kernel = nn.Parameter(w.contiguous())
        self.bias = nn.Parameter(b.contiguous())
        self._mask_pad()

    def forward(self, inputs, **kwargs):
        if isinstance(self.kernel, nn.UninitializedParameter):
            self.build(inputs.shape[-1], inputs.device)
        return ops.cross(inputs, self.kernel, self.bias).unsqueeze(-1)


# --------------------------------------------------------------------------------------
# a8  CIN (IL:285-327)
# --------------------------------------------------------------------------------------
class CIN(nn.Module):
    """IL:296.  ``[B,m,D] -> [B,1]`` (``output_dim == 1``: Dense(1) over the concatenated
    per-layer pools) or ``[B, n_layers*D]``.  ``precision``: ``"fp32"`` (CUDA-core parity
    mode, 1e-5) or ``"bf16"`` (tcgen05 tensor cores, fp32 accumulate, 2e-2)."""

    def __init__(self, conv_size=None, output_dim=1, precision="auto", seed=2020):
        super().__init__()
        self.conv_size = list(conv_size) if conv_size is not None else [200, 200, 200]
        self.output_dim = output_dim
        # "auto" (default): the bf16 tcgen05 path where its kernels cover the shape (26 fields, embedding
        # dim a multiple of 16, layer sizes <= 208), else the fp32 kernels -- so a default-constructed layer
        # works for every feature spec the reference accepts (e.g. sparse_fea_deal's cross_unit=8).
        # "bf16" / "fp32" force one path (bf16 raises KonError on shapes it does not cover).
        if precision not in ("auto", "fp32", "bf16"):
            raise ValueError("CIN precision must be 'auto', 'fp32' or 'bf16'")
        self.precision_mode = precision
        self.precision = {"fp32": L.KON_CIN_FP32, "bf16": L.KON_CIN_BF16, "auto": None}[precision]
        self.seed = seed
        self.conv_kernels = nn.ParameterList()             # Keras Conv1D kernels [1,C,N]
        self.conv_biases = nn.ParameterList()
        self.logit_kernel = None
        self.logit_bias = None

    def build(self, m: int, D: int, device):
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        hp = m
        for n in self.conv_size:
            c = hp * m
            lim = (6.0 / (c + n)) ** 0.5
            self.conv_kernels.append(nn.Parameter(((torch.rand(1, c, n, generator=g) * 2 - 1) * lim).to(device)))
            self.conv_biases.append(nn.Parameter(torch.zeros(n, device=device)))
            hp = n
        if self.output_dim == 1:
            k = len(self.conv_size) * D
            lim = (6.0 / (k + 1)) ** 0.5
            self.logit_kernel = nn.Parameter(((torch.rand(k, 1, generator=g) * 2 - 1) * lim).to(device))
            self.logit_bias = nn.Parameter(torch.zeros(1, device=device))

    def load_reference_weights(self, conv_kernels, conv_biases, logit_kernel=None, logit_bias=None):
        self.conv_kernels = nn.ParameterList([nn.Parameter(k.clone()) for k in conv_kernels])
        self.conv_biases = nn.ParameterList([nn.Parameter(b.clone()) for b in conv_biases])
        if logit_kernel is not None:
            self.logit_kernel = nn.Parameter(logit_kernel.clone())
            self.logit_bias = nn.Parameter(logit_bias.clone())

    def forward(self, inputs, fields=None, **kwargs):
        """``inputs``: ``[B,m,D]``; or the concat buffer ``[B,W]`` with ``fields=(m,D)`` (its first
        m*D columns are read in place and the gradient comes back as one ``[B,W]`` tensor)."""
        m, D = fields if fields is not None else (inputs.shape[1], inputs.shape[2])
        if len(self.conv_kernels) == 0:
            self.build(m, D, inputs.device)
        if self.precision is None:
            sizes = [int(k.shape[-1]) for k in self.conv_kernels]
            ok = m == 26 and D % 16 == 0 and all(1 <= n <= 208 for n in sizes)
            self.precision = L.KON_CIN_BF16 if ok else L.KON_CIN_FP32
        pooled = ops.cin(inputs, [k[0] for k in self.conv_kernels], list(self.conv_biases), self.precision,
                         fields=fields)
        if self.output_dim == 1:
            return torch.addmm(self.logit_bias, pooled, self.logit_kernel)      # IL:325
        return pooled


# --------------------------------------------------------------------------------------
# a9-a10  ProductAttentionLayer (BL:272-311), MultHeadAttentionLayer (BL:313-380)
# --------------------------------------------------------------------------------------
class ProductAttentionLayer(nn.Module):
    """BL:278.  ``[q, k, v]`` (each ``[..., F, d]``) -> ``sigmoid(mask(q k^T [/ sqrt d])) v`` (the attribute the
    reference calls ``softmax`` is a sigmoid, BL:286).  ``mask_mod`` 1: ``score @ float(mask)``; 2:
    ``score + float(mask) * -1e5`` (BL:299-306); ``mask`` is a ``[F,F]`` tensor shared by all samples (what
    SeqFM builds, MD:282-289)."""

    def __init__(self, use_scale=False, supports_masking=True, mask_mod=1):
        super().__init__()
        self.use_scale, self.supports_masking, self.mask_mod = use_scale, supports_masking, mask_mod

    def forward(self, inputs, mask=None, **kwargs):
        q, k, v = inputs
        return ops.product_attention(q, k, v, mask=mask, use_scale=self.use_scale, mask_mode=self.mask_mod)


class MultHeadAttentionLayer(nn.Module):
    """BL:318.  ``x [B,F,k_in]`` -> ``[atten_v, res]``, both ``[H,B,F,d]`` (BL:377).
    ``attention_head_dim`` is the number of heads H, ``attention_dim`` the per-head width d
    (BL:337-353).  ``value_w`` exists for weight-file compatibility but, as in the reference
    (BL:360), is never read: V = X key_w.

    ``forward`` returns the reference's pair; ``fused_block`` returns what the wrapping
    ``DnnLayer`` makes of it, ``ReLU(res + atten_v)`` (CL:205-216), in ONE kernel."""

    def __init__(self, attention_dim, attention_head_dim, seed=2020, use_scale=True, use_res=True,
                 use_ln=True, head_concat=False, supports_masking=True, atten_mask_mod=1, precision="fp32"):
        super().__init__()
        self.bf16 = {"fp32": False, "bf16": True}[precision]     # bf16: tensor-core path, tolerance 2e-2
        self.attention_dim, self.attention_head_dim = attention_dim, attention_head_dim
        self.seed, self.use_scale, self.use_res, self.use_ln = seed, use_scale, use_res, use_ln
        self.head_concat, self.atten_mask_mod = head_concat, atten_mask_mod
        self.attention_cal = ProductAttentionLayer(use_scale=use_scale, mask_mod=atten_mask_mod)   # BL:327
        self.query_w = nn.UninitializedParameter()

    def build(self, k_in: int, device):
        H, d = self.attention_head_dim, self.attention_dim
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        lim = (6.0 / (k_in * d + H * d)) ** 0.5            # Keras fans of a [k_in,H,d] tensor
        mk = lambda: nn.Parameter(((torch.rand(k_in, H, d, generator=g) * 2 - 1) * lim).to(device))
        self.query_w, self.key_w, self.value_w = mk(), mk(), mk()
        self.res_w = mk() if self.use_res else None
        if self.use_ln:
            self.ln_gamma = nn.Parameter(torch.ones(d, device=device))
            self.ln_beta = nn.Parameter(torch.zeros(d, device=device))
        else:
            self.ln_gamma = self.ln_beta = None

    def load_reference_weights(self, query_w, key_w, res_w=None, ln_gamma=None, ln_beta=None, value_w=None):
        self.query_w, self.key_w = nn.Parameter(query_w.clone()), nn.Parameter(key_w.clone())
        self.value_w = nn.Parameter((value_w if value_w is not None else key_w).clone())
        self.res_w = None if res_w is None else nn.Parameter(res_w.clone())
        self.ln_gamma = None if ln_gamma is None else nn.Parameter(ln_gamma.clone())
        self.ln_beta = None if ln_beta is None else nn.Parameter(ln_beta.clone())

    def _ensure(self, x):
        if isinstance(self.query_w, nn.UninitializedParameter):
            self.build(x.shape[-1], x.device)

    def fused_block(self, x, layout: str = "hbfd") -> torch.Tensor:
        """``ReLU(LN(sigmoid(QK^T/sqrt d) K) + X res_w)`` -> ``[H,B,F,d]`` (``layout``: memory order of the
        result on the bf16 path, ``ops._ATTN_LAYOUTS``)."""
        self._ensure(x)
        return ops.attention(x, self.query_w, self.key_w, self.res_w, self.ln_gamma, self.ln_beta,
                             use_scale=self.use_scale, use_ln=self.use_ln, use_res=self.use_res, relu=True,
                             bf16=self.bf16, layout=layout)

    def _forward_masked(self, inputs, mask):
        """BL:358-377 step by step for a masked call (SeqFM / BST reuse, MD:292-301): the three projections are
        cuBLAS GEMMs (``tensordot``), the masked product attention is ``kon_pattn_fwd/bwd``."""
        q = torch.tensordot(inputs, self.query_w, dims=1).permute(2, 0, 1, 3).contiguous()
        k = torch.tensordot(inputs, self.key_w, dims=1).permute(2, 0, 1, 3).contiguous()
        atten_v = self.attention_cal([q, k, k], mask=mask)          # v = X key_w (BL:360)
        if self.use_ln:
            atten_v = torch.nn.functional.layer_norm(atten_v, (self.attention_dim,), self.ln_gamma, self.ln_beta, 1e-3)
        if self.head_concat:
            atten_v = atten_v.permute(1, 0, 2, 3)
        if self.attention_head_dim == 1:
            return atten_v.squeeze(0)
        res = []
        if self.use_res:
            res = torch.tensordot(inputs, self.res_w, dims=1).permute(2, 0, 1, 3)
        return [atten_v, res]

    def forward(self, inputs, mask=None, **kwargs):
        self._ensure(inputs)
        if mask is not None:
            return self._forward_masked(inputs, mask)
        atten_v = ops.attention(inputs, self.query_w, self.key_w, None, self.ln_gamma, self.ln_beta,
                                use_scale=self.use_scale, use_ln=self.use_ln, use_res=False, relu=False)
        if self.head_concat:
            atten_v = atten_v.permute(1, 0, 2, 3)
        if self.attention_head_dim == 1:
            return atten_v.squeeze(0)                      # BL:374-375
        res = []
        if self.use_res:
            res = torch.tensordot(inputs, self.res_w, dims=1).permute(2, 0, 1, 3)   # BL:366
        return [atten_v, res]


class AttentionBaseLayer(nn.Module):
    """IL:335-366, AFM's pooling over the pairwise products.  In the reference the attention weight is
    ``Activation('softmax')`` of a ``[B,P,1]`` score (IL:340-341, 362-363): the softmax runs over the LAST
    axis, whose size is 1, so every weight is exactly 1 and the layer computes
    ``Dense(output_dim)(sum_p pair_p)``; the scoring weights (``single_score_w/b``, the ``Dense(1,relu)``
    kernel) exist, keep their reference shapes for weight files, and receive no gradient -- the fixture from
    the reference's own source (tests/golden/ref_models.npz, afm) shows exactly zeros there.
    ``forward`` takes the reference's list of products; ``fused(v)`` takes the field embeddings and uses
    ``sum_{i<j} v_i v_j`` from the FM kernel without materialising the ``[B, F(F-1)/2, k]`` tensor."""

    def __init__(self, attention_dim=4, seed=2020, output_dim=1):
        super().__init__()
        self.atten_dim, self.seed, self.output_dim = attention_dim, seed, output_dim
        self.kernel_w = nn.UninitializedParameter()

    def build(self, k: int, device):
        g = torch.Generator(device="cpu").manual_seed(self.seed)

        def glorot(*shape):
            fi, fo = (shape[0], shape[0]) if len(shape) == 1 else shape
            lim = (6.0 / (fi + fo)) ** 0.5
            return nn.Parameter(((torch.rand(*shape, generator=g) * 2 - 1) * lim).to(device))
        self.kernel_w, self.kernel_b = glorot(k, self.atten_dim), glorot(self.atten_dim)
        self.mlp_kernel = glorot(self.atten_dim, 1)
        self.out_kernel = glorot(k, self.output_dim)
        self.out_bias = nn.Parameter(torch.zeros(self.output_dim, device=device))

    def load_reference_weights(self, score_w, score_b, mlp_w, out_w, out_b):
        self.kernel_w, self.kernel_b = nn.Parameter(score_w.clone()), nn.Parameter(score_b.clone())
        self.mlp_kernel = nn.Parameter(mlp_w.clone())
        self.out_kernel, self.out_bias = nn.Parameter(out_w.clone()), nn.Parameter(out_b.clone())

    def _dense(self, pooled):
        if isinstance(self.kernel_w, nn.UninitializedParameter):
            self.build(pooled.shape[-1], pooled.device)
        if ops.head_supported(pooled, None, self.out_kernel):
            return ops.head(pooled, None, self.out_kernel, self.out_bias)
        return torch.addmm(self.out_bias, pooled, self.out_kernel)

    def fused(self, v: torch.Tensor) -> torch.Tensor:
        return self._dense(ops.fm(v, None))

    def forward(self, inputs, **kwargs):
        x = pack_fields(inputs)                            # tf.concat(inputs, axis=1)  [B,P,k]
        return self._dense(ops.pool_sum(x))                # reduce_sum(1 * inputs, axis=1) -> Dense


# --------------------------------------------------------------------------------------
# a11-a12 glue, MLP, heads (CL) -- torch / cuBLAS, adjacent to the hot path
# --------------------------------------------------------------------------------------
class StackLayer(nn.Module):
    """CL:32-55: Flatten each input and concat on ``axis`` (default -1)."""

    def __init__(self, use_flat=True, axis=None):
        super().__init__()
        self.use_flat, self.axis = use_flat, axis

    def forward(self, inputs, **kwargs):
        if isinstance(inputs, torch.Tensor):
            inputs = [inputs]
        p = getattr(inputs, "packed", None)
        if p is not None:                                  # all fields of one packed buffer
            return p.reshape(p.shape[0], -1) if self.use_flat else p
        if self.use_flat:
            inputs = [t.reshape(t.shape[0], -1) for t in inputs]
        if len(inputs) == 1:
            return inputs[0]
        return torch.cat(list(inputs), dim=self.axis if self.axis else -1)


def _bias_act_gemm(x, w, b, relu: bool):
    """``x @ w + b`` with the bias (and ReLU) applied in the GEMM epilogue (cuBLASLt), one kernel."""
    if relu:
        return torch._addmm_activation(b, x, w)            # epilogue: bias + ReLU
    return torch.addmm(b, x, w)


def _relu_grad(g, y):
    return torch.ops.aten.threshold_backward(g, y, 0)     # g * (y > 0), one kernel, from the saved OUTPUT


class _DenseFn(torch.autograd.Function):
    """``[ReLU](x @ w + b)`` (CL:190 + CL:216 when no residual sits in between): bias and activation live in the GEMM
    epilogue; the bias gradient is a GEMV ``ones^T @ g`` on cuBLAS instead of torch's column reduction
    (65,536 x 256 bf16: 65 us -> a few us)."""

    @staticmethod
    def forward(ctx, x, w, b, relu=False):
        y = _bias_act_gemm(x, w, b, relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.relu = relu
        ctx.x_key = ops.xgrad_key(x)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, y = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = _relu_grad(g, y)
        gx = g @ w.t() if ctx.needs_input_grad[0] else None
        ops.offer_xgrad(ctx.x_key, gx)      # a later branch on the same input adds its gradient into gx (ops._XGRAD)
        gw = x.t() @ g if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            ones = torch.ones((1, g.shape[0]), dtype=g.dtype, device=g.device)
            gb = (ones @ g).reshape(-1)
        return gx, gw, gb, None


def _mm_f32_out(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 GEMM (fp32 accumulate, the result is never rounded to bf16)."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.mm(a, b).float()


class _DenseCastFn(torch.autograd.Function):
    """First MLP layer on a reduced-precision path: ``x`` arrives in fp32 (the concat buffer), is cast to the
    compute dtype here, and its gradient goes back in fp32 straight out of the GEMM (fp32 accumulators written
    as fp32) -- no bf16 rounding of the gradient that flows into the embedding rows, and no separate
    bf16 -> fp32 cast pass over the [B, 13+F*k] gradient."""

    @staticmethod
    def forward(ctx, x, w, b, relu=False):
        xc = x.to(w.dtype)
        y = _bias_act_gemm(xc, w, b, relu)
        ctx.save_for_backward(xc, w, y if relu else None)
        ctx.relu = relu
        ctx.x_key = ops.xgrad_key(x)
        return y

    @staticmethod
    def backward(ctx, g):
        xc, w, y = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = _relu_grad(g, y)
        gx = _mm_f32_out(g, w.t()) if ctx.needs_input_grad[0] else None
        ops.offer_xgrad(ctx.x_key, gx)
        gw = xc.t() @ g if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            ones = torch.ones((1, g.shape[0]), dtype=g.dtype, device=g.device)
            gb = (ones @ g).reshape(-1)
        return gx, gw, gb, None


class _DenseMasterFn(torch.autograd.Function):
    """Reduced-precision Dense layer over fp32 MASTER weights: ``w`` / ``b`` are the fp32 parameters (their gradients
    leave the weight-gradient GEMM and the bias GEMV as fp32 accumulators, no bf16 gradient and no bf16 -> fp32 cast
    pass per parameter), ``wc`` / ``bc`` their compute-dtype copies (made by the layer for all its weights in one
    multi-tensor launch).  A fp32 ``x`` (the concat buffer) is cast here and gets its gradient back in fp32, straight
    out of the dgrad GEMM, like ``_DenseCastFn``."""

    @staticmethod
    def forward(ctx, x, w, b, wc, bc, relu=False):
        ctx.x_f32 = x.dtype == torch.float32
        ctx.x_key = ops.xgrad_key(x)
        xc = x.to(wc.dtype) if x.dtype != wc.dtype else x
        y = _bias_act_gemm(xc, wc, bc, relu)
        ctx.save_for_backward(xc, wc, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, g):
        xc, wc, y = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = _relu_grad(g, y)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _mm_f32_out(g, wc.t()) if ctx.x_f32 else g @ wc.t()
            if ctx.x_f32:
                ops.offer_xgrad(ctx.x_key, gx)
        gw = _mm_f32_out(xc.t(), g) if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            ones = torch.ones((1, g.shape[0]), dtype=g.dtype, device=g.device)
            gb = _mm_f32_out(ones, g).reshape(-1)
        return gx, gw, gb, None, None, None


class DnnLayer(nn.Module):
    """CL:159-226 with the defaults the CTR builders use (``res_unit=1``, no BN/LN, ReLU):
    per hidden layer ``Dense`` -> ``Add([ori, x])`` when the shapes allow it (CL:206-214)
    -> ReLU; optional ``Dense(output_dim)`` logit layer.  ``other_dense=[layer]`` wraps a
    MultHeadAttentionLayer as the hidden layer (AutoInt, MD:160-161)."""

    def __init__(self, hidden_units=None, hidden_activate=None, use_bn=False, res_unit=1, output_dim=-1,
                 seed=2020, other_dense=None, use_ln=False, use_flatten=False, **kwargs):
        super().__init__()
        if use_bn or use_ln:
            raise L.KonError("DnnLayer(use_bn/use_ln) is outside the CTR hot path")
        self.hidden_units = list(hidden_units) if hidden_units is not None else []
        self.output_dim, self.seed = output_dim, seed
        self.other_dense = nn.ModuleList(other_dense) if other_dense else None
        self.kernels = nn.ParameterList()
        self.biases = nn.ParameterList()
        self.logit_kernel = self.logit_bias = None
        self.compute_dtype = None                          # e.g. torch.bfloat16 for the bench

    def build(self, in_dim: int, device):
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        d = in_dim
        for u in self.hidden_units:
            lim = (6.0 / (d + u)) ** 0.5
            self.kernels.append(nn.Parameter(((torch.rand(d, u, generator=g) * 2 - 1) * lim).to(device)))
            blim = (6.0 / (u + u)) ** 0.5                  # bias_initializer=glorot (CL:190)
            self.biases.append(nn.Parameter(((torch.rand(u, generator=g) * 2 - 1) * blim).to(device)))
            d = u
        if self.output_dim != -1:
            lim = (6.0 / (d + self.output_dim)) ** 0.5
            self.logit_kernel = nn.Parameter(((torch.rand(d, self.output_dim, generator=g) * 2 - 1) * lim).to(device))
            self.logit_bias = nn.Parameter(torch.zeros(self.output_dim, device=device))

    def _shadow_weights(self, cd):
        """Compute-dtype copies of all kernels / biases, refreshed from the fp32 masters in ONE multi-tensor launch
        per forward (the per-parameter ``w.to(bf16)`` were 2 small launches per layer, and autograd cast every
        bf16 gradient back to fp32 with another one)."""
        params = [t for wb in zip(self.kernels, self.biases) for t in wb]
        sh = getattr(self, "_shadow", None)
        if sh is None or len(sh) != len(params) or sh[0].dtype != cd or any(
                a.shape != b.shape or a.device != b.device for a, b in zip(sh, params)):
            sh = self._shadow = [torch.empty_like(t, dtype=cd) for t in params]
        with torch.no_grad():
            torch._foreach_copy_(sh, [t.detach() for t in params])
        return sh

    def load_reference_weights(self, kernels, biases, logit_kernel=None, logit_bias=None):
        self.kernels = nn.ParameterList([nn.Parameter(k.clone()) for k in kernels])
        self.biases = nn.ParameterList([nn.Parameter(b.clone()) for b in biases])
        if logit_kernel is not None:
            self.logit_kernel, self.logit_bias = nn.Parameter(logit_kernel.clone()), nn.Parameter(logit_bias.clone())

    def forward(self, x, **kwargs):
        if self.other_dense is not None:                   # AutoInt wiring
            for layer in self.other_dense:
                x = layer.fused_block(x, **({"layout": kwargs["layout"]} if "layout" in kwargs else {}))
            return x
        if len(self.kernels) == 0 and self.hidden_units:
            self.build(x.shape[-1], x.device)
        cd = self.compute_dtype
        shadow = self._shadow_weights(cd) if (cd is not None and cd != torch.float32 and x.is_cuda) else None
        for i, (w, b) in enumerate(zip(self.kernels, self.biases)):
            ori = x
            res = w.shape[0] == w.shape[1]                 # the reference's Add([ori, x]) fires (CL:206-214)
            fuse = not res                                 # else ReLU follows the residual add, outside the GEMM
            if shadow is not None:
                x = _DenseMasterFn.apply(x, w, b, shadow[2 * i], shadow[2 * i + 1], fuse)
            elif cd is not None and x.dtype != cd:
                x = _DenseCastFn.apply(x, w.to(cd), b.to(cd), fuse) if i == 0 else _DenseFn.apply(x.to(cd), w.to(cd), b.to(cd), fuse)
            else:
                x = _DenseFn.apply(x, w.to(x.dtype), b.to(x.dtype), fuse)
            if res:
                x = torch.relu(ori.to(x.dtype) + x)
        if cd is not None and x.dtype != cd:      # no hidden layer: the logit layer still runs in the compute dtype
            x = x.to(cd)
        if self.logit_kernel is not None:
            x = torch.addmm(self.logit_bias.to(x.dtype), x, self.logit_kernel.to(x.dtype))
        return x.float() if cd is not None else x


class MergeScoreLayer(nn.Module):
    """CL:86-100: flatten + concat, then ``Dense(2, softmax)``."""

    def __init__(self, use_merge: bool = True, output_dim=2, seed=2020):
        super().__init__()
        self.use_merge, self.output_dim, self.seed = use_merge, output_dim, seed
        self.kernel = nn.UninitializedParameter()
        self.bias = None

    def build(self, d, device):
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        lim = (6.0 / (d + self.output_dim)) ** 0.5
        self.kernel = nn.Parameter(((torch.rand(d, self.output_dim, generator=g) * 2 - 1) * lim).to(device))
        self.bias = nn.Parameter(torch.zeros(self.output_dim, device=device))

    def load_reference_weights(self, kernel, bias):
        self.kernel, self.bias = nn.Parameter(kernel.clone()), nn.Parameter(bias.clone())

    def logits(self, inputs):
        parts = [t.reshape(t.shape[0], -1) for t in inputs] if self.use_merge else [inputs.reshape(inputs.shape[0], -1)]
        d = sum(t.shape[1] for t in parts)
        if isinstance(self.kernel, nn.UninitializedParameter):
            self.build(d, parts[0].device)
        x1, x2 = parts[0], (parts[1] if len(parts) == 2 else None)
        if len(parts) <= 2 and x1.is_cuda and ops.head_supported(x1, x2, self.kernel):
            return ops.head(x1, x2, self.kernel, self.bias)        # both inputs read in place: no Concatenate copy
        return torch.addmm(self.bias, torch.cat(parts, dim=-1) if len(parts) > 1 else x1, self.kernel)

    def forward(self, inputs, **kwargs):
        return torch.softmax(self.logits(inputs), dim=-1)


class ScoreLayer(nn.Module):
    """CL:58-84 (``use_inner``/``use_global`` off): optional Keras Add, then sigmoid."""

    def __init__(self, use_add=False, **kwargs):
        super().__init__()
        self.use_add = use_add

    @staticmethod
    def summed(inputs):
        # Keras Add: lower-rank inputs are expanded at axis 1 until all ranks match (_Merge.call)
        nd = max(t.dim() for t in inputs)
        inputs = [t.reshape(tuple(t.shape[:1]) + (1,) * (nd - t.dim()) + tuple(t.shape[1:])) for t in inputs]
        out = inputs[0]
        for t in inputs[1:]:
            out = out + t
        return out

    def forward(self, inputs, **kwargs):
        if self.use_add:
            inputs = self.summed(inputs)
        return torch.sigmoid(inputs)
