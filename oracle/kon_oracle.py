"""CPU oracle for the CTR hot path of TIXhjq/ML_Function (``kon.model.ctr_model``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``ml_function_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the thing shipped.

PARITY PINNED TO THE REFERENCE'S OWN SOURCE (round 2).  The reference holds no tests,
golden vectors or fixtures, and its arithmetic lives in TensorFlow 2.1, which is not
installable here.  But its layer files are plain Python: ``oracle/run_reference.py``
imports the UNMODIFIED classes from /root/reference over a minimal eager ``tensorflow``
shim (``oracle/_ref_shim``, torch-CPU) and ``tests/golden/make_ref_golden.py`` executes
them -- every layer of the hot path and the builders FM / DeepFM / DCN / XDeepFM /
AutoInt / NFM / AFM end to end, forward, loss and all gradients -- into the committed
fixtures ``tests/golden/ref_layers.npz`` / ``ref_models.npz``.
``tests/test_ref_pinned_cpu.py`` asserts that the functions below reproduce those
fixtures BIT FOR BIT in fp32, and (where /root/reference exists) that the fixtures are
what the reference produces today.  What remains assumed is only what the shim itself
supplies -- the Keras primitives listed in ``oracle/_ref_shim/README.md`` (Embedding =
cast + gather, Add = left-to-right sum with rank expansion at axis 1, Dense / Conv1D(1)
= matmul + bias, LayerNormalization eps 1e-3, K.dot / K.batch_dot) -- written from the
published TF-2.1 behaviour; a running TensorFlow is still not available to check those.

Short names for citations (paths under the reference repo):
  IL = kon/model/ctr_model/layer/interactive_layer/interactive_layer.py
  CL = kon/model/ctr_model/layer/core_layer/core_layer.py
  BL = kon/model/ctr_model/layer/behavior_layer/behavior_layer.py
  MD = kon/model/ctr_model/model/models.py
  DP = kon/utils/data_prepare.py

Every function works in the dtype of its inputs (fp32 = the reference's
arithmetic, fp64 = "truth" for judging fp32 error) and is differentiable by
torch autograd, which stands in for ``tf.GradientTape`` inside ``Model.fit``.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence

import numpy as np
import torch

# --------------------------------------------------------------------------- #
# Keras primitives the reference leans on
# --------------------------------------------------------------------------- #


def keras_add(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """``tf.keras.layers.Add``: left-to-right sequential sum with broadcasting
    of size-1 dims (TF ``_Merge._merge_function``: ``out = x[0]; out += x[i]``).
    Inputs of lower rank are first expanded at axis 1 until all ranks match
    (``_Merge.call``: "expand each of them at axis=1"), e.g. ``[B,1,1] + [B,1]``
    is ``[B,1,1]``, not numpy's right-aligned ``[B,B,1]``."""
    nd = max(t.dim() for t in tensors)
    tensors = [t.reshape(tuple(t.shape[:1]) + (1,) * (nd - t.dim()) + tuple(t.shape[1:])) for t in tensors]
    out = tensors[0]
    for t in tensors[1:]:
        out = out + t
    return out


def keras_flatten(x: torch.Tensor) -> torch.Tensor:
    """``tf.keras.layers.Flatten``: ``[B, ...] -> [B, prod(...)]``."""
    return x.reshape(x.shape[0], -1)


def keras_dense(x, kernel, bias=None, activation=None):
    """``tf.keras.layers.Dense``: ``x @ kernel + bias`` on the last axis;
    kernel is ``[in, units]`` (Keras layout)."""
    y = torch.matmul(x, kernel)
    if bias is not None:
        y = y + bias
    if activation == "softmax":
        y = torch.softmax(y, dim=-1)
    elif activation == "relu":
        y = torch.relu(y)
    elif activation == "sigmoid":
        y = torch.sigmoid(y)
    return y


def keras_layer_norm(x, gamma, beta, eps: float = 1e-3):
    """``tf.keras.layers.LayerNormalization()`` defaults: axis=-1, epsilon=1e-3,
    biased variance; TF computes ``nn.moments`` then ``nn.batch_normalization``:
    ``inv = rsqrt(var+eps)*gamma ; y = x*inv + (beta - mean*inv)``."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    inv = torch.rsqrt(var + eps) * gamma
    return x * inv + (beta - mean * inv)


def keras_conv1d_k1(x, kernel, bias):
    """``tf.keras.layers.Conv1D(n, 1)``: kernel ``[1, C_in, n]``, bias ``[n]``,
    linear activation, channels-last: per-position Dense."""
    return torch.matmul(x, kernel[0]) + bias


# --------------------------------------------------------------------------- #
# a1-a3: embeddings and pooling
# --------------------------------------------------------------------------- #


def sparse_embed(ids: Sequence[torch.Tensor], tables: Sequence[torch.Tensor],
                 use_flatten: bool = True, use_add: bool = False):
    """``SparseEmbed.call`` (IL:225-242).  ``ids[f]`` is ``[B, input_length]``
    (float32 in the reference, DP:290-292; Keras ``Embedding`` casts to int32
    and gathers), ``tables[f]`` is ``[word_size_f, dim]``.  Returns the list of
    ``[B, input_length, dim]`` gathers, each Flatten-ed when ``use_flatten``
    (IL:230-231), then ``Add``-ed over fields when ``use_add`` (IL:233-234)."""
    embed_list = [t[i.to(torch.int64)] for t, i in zip(tables, ids)]  # IL:227-228
    if use_flatten:
        embed_list = [keras_flatten(e) for e in embed_list]
    if use_add:
        embed_list = keras_add(embed_list)
    return embed_list


def seq_base_layer(inputs: Sequence[torch.Tensor]):
    """``SeqBaseLayer.call`` (BL:45-46): sum-pool each ``[B, L, k]`` over L.
    Padding id 0 still contributes row 0 -- the mask is only forwarded
    (BL:48-51)."""
    return [torch.sum(x, dim=1).unsqueeze(1) for x in inputs]


def embedding_grad(ids: np.ndarray, d_out: np.ndarray, n_rows: int):
    """Embedding backward for ONE table (implicit in ``Model.fit``; TF emits an
    IndexedSlices that the optimizer de-duplicates with ``unique`` +
    ``unsorted_segment_sum``).  ``ids`` ``[N]`` ints, ``d_out`` ``[N, k]``.
    Returns ``(unique_rows ascending [U], grads [U, k])`` where each row's
    gradient is summed in input order -- ``np.add.at`` is unbuffered and
    sequential, so the fp32 rounding order is "ascending sample index"."""
    ids = np.asarray(ids).astype(np.int64).reshape(-1)
    d_out = np.asarray(d_out).reshape(ids.shape[0], -1)
    assert ids.size == 0 or (ids.min() >= 0 and ids.max() < n_rows)
    uniq, inv = np.unique(ids, return_inverse=True)
    grads = np.zeros((uniq.shape[0], d_out.shape[1]), dtype=d_out.dtype)
    np.add.at(grads, inv, d_out)
    return uniq, grads


# --------------------------------------------------------------------------- #
# a5-a6: FM
# --------------------------------------------------------------------------- #


def inner_layer(inputs: Sequence[torch.Tensor], use_add: bool = False):
    """``InnerLayer.call`` with ``use_inner=True`` (IL:59-66): Hadamard product
    of every pair in ``itertools.combinations`` order, optionally Keras-Add-ed
    left to right."""
    cross_list = [a * b for a, b in itertools.combinations(inputs, 2)]
    if use_add:
        cross_list = keras_add(cross_list)
    return cross_list


def fm_layer(cross_embed: Sequence[torch.Tensor], linear_embed: Sequence[torch.Tensor]):
    """``FmLayer.call`` (IL:161-170) with the default ``use_add=True``:
    ``Add([Add(pairwise products)] + linear_list)``; ``[B,1,k] + [B,1,1]``
    broadcasts so the result keeps the embedding axis."""
    cross = inner_layer(cross_embed, use_add=True)
    return keras_add([cross] + list(linear_embed))


def fm_closed_form(v: torch.Tensor, lin: torch.Tensor):
    """Closed form of fm_layer on packed inputs ``v [B,F,k]``, ``lin [B,F]``:
    ``0.5((sum v)^2 - sum v^2) + sum lin`` -> ``[B,k]``."""
    s = v.sum(dim=1)
    q = (v * v).sum(dim=1)
    return 0.5 * (s * s - q) + lin.sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------- #
# a7: DCN cross
# --------------------------------------------------------------------------- #


def cross_layer(x: torch.Tensor, kernels: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
    """``CrossLayer.call`` (IL:275-282).  ``x [B,D]``; ``kernels[i]``,
    ``biases[i]`` are ``[D,1]``.  Returns ``[B,D,1]``:
    ``pre = batch_dot(x0, dot(pre^T, w_i)) + pre + b_i``."""
    inputs = x.unsqueeze(-1)                                      # IL:276
    pre = inputs
    for w, b in zip(kernels, biases):
        s = torch.matmul(pre.transpose(1, 2), w)                  # K.dot -> [B,1,1]
        pre = torch.matmul(inputs, s) + pre + b                   # K.batch_dot + adds
    return pre


# --------------------------------------------------------------------------- #
# a8: xDeepFM CIN
# --------------------------------------------------------------------------- #


def cin(inputs: torch.Tensor, conv_kernels: Sequence[torch.Tensor],
        conv_biases: Sequence[torch.Tensor], logit_kernel: Optional[torch.Tensor] = None,
        logit_bias: Optional[torch.Tensor] = None, return_pooled: bool = False):
    """``CIN.call`` (IL:310-327).  ``inputs [B,m,D]``; ``conv_kernels[l]`` is the
    Keras Conv1D kernel ``[1, H_{l-1}*m, H_l]`` (channel index ``h*m + i``,
    from the transpose ``[1,0,3,2]`` + reshape at IL:317-318), bias ``[H_l]``.
    No activation; pooling is over the feature-map axis (IL:322) so each layer
    contributes ``[B,D]``; ``output_dim==1`` applies ``Dense(1)`` (IL:304,325)."""
    x0 = list(torch.split(inputs, 1, dim=-1))                     # IL:311  D x [B,m,1]
    pre = x0
    pooled = []
    for kern, bias in zip(conv_kernels, conv_biases):
        # tf.matmul packs its python-list operands afresh on every iteration (IL:316), so x0's
        # gradient reaches `inputs` as one contribution per layer, summed in layer order
        z = torch.matmul(torch.stack(x0, dim=0), torch.stack(pre, dim=0).transpose(-1, -2))   # [D,B,m,H]
        z = z.permute(1, 0, 3, 2)                                 # IL:317  [B,D,H,m]
        z = z.reshape(-1, z.shape[1], z.shape[2] * z.shape[3])    # IL:318  [B,D,H*m]
        z = keras_conv1d_k1(z, kern, bias)                        # IL:319  [B,D,N]
        pre = z.transpose(1, 2)                                   # IL:320  [B,N,D]
        pre = list(torch.split(pre, 1, dim=-1))                   # IL:321  D x [B,N,1]
        pooled.append(torch.sum(z, dim=-1))                       # IL:322  [B,D]
    output = torch.cat(pooled, dim=-1)                            # IL:323
    if return_pooled:
        return output
    if logit_kernel is not None:
        output = keras_dense(output, logit_kernel, logit_bias)    # IL:325
    return output


def cin_closed_form(x0: torch.Tensor, conv_kernels, conv_biases):
    """Closed form used by the kernels: per layer
    ``z[b,d,o] = sum_{h,i} W[h*m+i,o] pre[b,h,d] x0[b,i,d] + bias[o]``,
    ``pre' = z^T``, ``pool[b,d] = sum_o z[b,d,o]``.  Returns pooled ``[B, L*D]``
    and the list of z ``[B,D,H_l]``."""
    B, m, D = x0.shape
    pre = x0
    pooled, zs = [], []
    for kern, bias in zip(conv_kernels, conv_biases):
        H = pre.shape[1]
        W = kern[0].reshape(H, m, -1)
        z = torch.einsum("bhd,bid,hio->bdo", pre, x0, W) + bias
        zs.append(z)
        pooled.append(z.sum(-1))
        pre = z.transpose(1, 2)
    return torch.cat(pooled, dim=-1), zs


# --------------------------------------------------------------------------- #
# a9-a10: AutoInt attention
# --------------------------------------------------------------------------- #


def product_attention(q, k, v, use_scale: bool = False, mask=None, mask_mod: int = 1):
    """``ProductAttentionLayer.call`` (BL:292-311): ``sigmoid`` (BL:286, the
    attribute is merely *named* softmax) of ``q k^T`` (``/ sqrt(d)`` when
    ``use_scale``), optional mask (mod 1: ``score @ mask``; mod 2:
    ``score + mask * -1e5``), times ``v``."""
    score = torch.matmul(q, k.transpose(-1, -2))
    if use_scale:
        score = score / (q.shape[-1] ** 0.5)
    if mask is not None:
        m = mask.to(score.dtype)
        if mask_mod == 1:
            score = torch.matmul(score, m)
        if mask_mod == 2:
            score = score + m * (-100000)
    score = torch.sigmoid(score)
    return torch.matmul(score, v)


def mult_head_attention(x, query_w, key_w, res_w=None, ln_gamma=None, ln_beta=None,
                        use_scale=True, use_res=True, use_ln=True, head_concat=False, mask=None,
                        atten_mask_mod=1):
    """``MultHeadAttentionLayer.call`` (BL:356-377).  ``x [B,F,k_in]``; weights
    ``[k_in, H, d]``.  ``v`` is projected with **key_w** (BL:360; ``value_w``
    is created at BL:346-349 but never read).  Returns ``[atten_v, res]`` with
    both ``[H,B,F,d]`` (or the squeezed ``atten_v`` when ``H == 1``, BL:374-375)."""
    H = query_w.shape[1]
    q = torch.tensordot(x, query_w, dims=1).permute(2, 0, 1, 3)   # BL:358
    k = torch.tensordot(x, key_w, dims=1).permute(2, 0, 1, 3)     # BL:359
    v = torch.tensordot(x, key_w, dims=1).permute(2, 0, 1, 3)     # BL:360
    atten_v = product_attention(q, k, v, use_scale=use_scale, mask=mask, mask_mod=atten_mask_mod)
    res = []
    if use_res:
        res = torch.tensordot(x, res_w, dims=1).permute(2, 0, 1, 3)  # BL:366
    if use_ln:
        atten_v = keras_layer_norm(atten_v, ln_gamma, ln_beta)    # BL:368-369
    if head_concat:
        atten_v = atten_v.permute(1, 0, 2, 3)                     # BL:371-372
    if H == 1:
        return atten_v.squeeze(0)                                 # BL:374-375
    return [atten_v, res]


def autoint_block(x, query_w, key_w, res_w, ln_gamma, ln_beta, use_scale=True, mask=None, atten_mask_mod=1):
    """What ``DnnLayer(res_unit=1, other_dense=[MultHeadAttentionLayer])`` does to
    ``x`` (CL:201-226 around BL:356-377): ``ReLU(Add([res, atten_v]))`` ->
    ``[H,B,F,d]``.  (``hidden_layer(x)`` returns ``[atten_v, res]`` which the
    loop unpacks as ``[x, ori]``; idx 0: ``res=[ori,x]``; ``Add`` fires since the
    shapes match; ``use_bn/use_ln`` of ResActivateLayer default False.)"""
    atten_v, res = mult_head_attention(x, query_w, key_w, res_w, ln_gamma, ln_beta,
                                       use_scale=use_scale, mask=mask, atten_mask_mod=atten_mask_mod)
    return torch.relu(keras_add([res, atten_v]))


def attention_base_layer(pairs: Sequence[torch.Tensor], score_w, score_b, mlp_w, out_w, out_b):
    """``AttentionBaseLayer.call`` (IL:359-366), AFM's pooling: ``concat(pairs, 1)`` ``[B,P,k]``;
    ``score = relu((x @ score_w + score_b) @ mlp_w)`` ``[B,P,1]`` (``Dense(1,'relu',use_bias=False)``,
    IL:340); ``Activation('softmax')`` normalises over the LAST axis, which has size 1, so every
    attention weight is exactly 1 (IL:341,363) and the scoring weights receive zero gradient;
    ``Dense(output_dim)(sum_p weight * x)``."""
    x = torch.cat(list(pairs), dim=1)
    score = torch.relu(torch.matmul(torch.matmul(x, score_w) + score_b, mlp_w))
    weight = torch.softmax(score, dim=-1)
    pooled = torch.sum(weight * x, dim=1)
    return keras_dense(pooled, out_w, out_b)


# --------------------------------------------------------------------------- #
# a11-a12: glue, MLP, heads
# --------------------------------------------------------------------------- #


def stack_layer(inputs: Sequence[torch.Tensor], use_flat: bool = True, axis: Optional[int] = None):
    """``StackLayer.call`` (CL:49-55): Flatten each, concat on ``axis``
    (default -1; note ``if axis:`` at CL:39 treats 0 like None)."""
    if use_flat:
        inputs = [keras_flatten(t) for t in inputs]
    if len(inputs) == 1:
        return inputs[0]
    return torch.cat(list(inputs), dim=axis if axis else -1)


def dnn_layer(x, kernels, biases, logit_kernel=None, logit_bias=None):
    """``DnnLayer.call`` (CL:201-226) with default ``res_unit=1``, ``use_bn``/
    ``use_ln`` False and Dense hidden layers: per layer ``Dense`` -> try
    ``Add([ori, x])`` (only possible when in/out dims match, else ValueError ->
    ``x``) -> ReLU; optional ``Dense(output_dim)`` logit layer (CL:223-224)."""
    for w, b in zip(kernels, biases):
        ori = x
        x = keras_dense(x, w, b)
        if ori.shape == x.shape:
            x = ori + x
        x = torch.relu(x)
    if logit_kernel is not None:
        x = keras_dense(x, logit_kernel, logit_bias)
    return x


def merge_score_layer(inputs, kernel, bias, use_merge: bool = True):
    """``MergeScoreLayer.call`` (CL:96-100): flatten+concat then
    ``Dense(2, softmax)``."""
    if use_merge:
        inputs = stack_layer(inputs)
    return keras_dense(inputs, kernel, bias, activation="softmax")


def score_layer(inputs, use_add: bool = False):
    """``ScoreLayer.call`` (CL:75-84) without ``use_inner``/``use_global``:
    optional Keras Add then sigmoid."""
    if use_add:
        inputs = keras_add(inputs)
    return torch.sigmoid(inputs)


def binary_crossentropy(y_true, y_pred, eps: float = 1e-7):
    """``tf.losses.binary_crossentropy`` on probabilities (EX un_seq.py:61):
    clip to ``[eps, 1-eps]``, mean over the last axis, then Keras averages
    over the batch."""
    p = torch.clamp(y_pred, eps, 1 - eps)
    bce = -(y_true * torch.log(p + eps) + (1 - y_true) * torch.log(1 - p + eps))
    return bce.mean(dim=-1).mean()


# --------------------------------------------------------------------------- #
# Model builders (forward only; autograd supplies the backward)
# --------------------------------------------------------------------------- #


class OracleParams(dict):
    """Plain name->tensor bag shared by the oracle models and the tests."""


def _embed_lists(p, sparse_ids):
    """FeatureInput (DP:65-76) with ``useFlattenSparse=False`` /
    ``useFlattenLinear=False``: lists of ``[B,1,k]`` and ``[B,1,1]``."""
    F = sparse_ids.shape[1]
    ids = [sparse_ids[:, f:f + 1] for f in range(F)]
    sparse = sparse_embed(ids, [p[f"emb_{f}"] for f in range(F)], use_flatten=False)
    linear = sparse_embed(ids, [p[f"lin_{f}"] for f in range(F)], use_flatten=False)
    return sparse, linear


def _dense_list(dense):
    return [dense[:, j:j + 1] for j in range(dense.shape[1])]


def model_fm(p, dense, sparse_ids):
    """``FM`` (MD:36-41)."""
    sparse, linear = _embed_lists(p, sparse_ids)
    fm_ = fm_layer(sparse, linear)
    return merge_score_layer(fm_.squeeze(1), p["head_w"], p["head_b"], use_merge=False)


def model_deepfm(p, dense, sparse_ids, n_hidden=3):
    """``DeepFM`` (MD:80-90)."""
    sparse, linear = _embed_lists(p, sparse_ids)
    fm_ = fm_layer(sparse, linear)
    dnn_in = stack_layer(_dense_list(dense) + sparse)
    dnn_ = dnn_layer(dnn_in, [p[f"dnn_w{i}"] for i in range(n_hidden)],
                     [p[f"dnn_b{i}"] for i in range(n_hidden)])
    return merge_score_layer([fm_, dnn_], p["head_w"], p["head_b"])


def model_dcn(p, dense, sparse_ids, cross_hidden=3, n_hidden=3):
    """``DCN`` (MD:92-106)."""
    sparse, _ = _embed_lists(p, sparse_ids)
    x0 = stack_layer(_dense_list(dense) + sparse)
    cross = cross_layer(x0, [p[f"outer_weight_{i}"] for i in range(cross_hidden)],
                        [p[f"outer_bias_{i}"] for i in range(cross_hidden)])
    deep = dnn_layer(x0, [p[f"dnn_w{i}"] for i in range(n_hidden)],
                     [p[f"dnn_b{i}"] for i in range(n_hidden)])
    return merge_score_layer([cross, deep], p["head_w"], p["head_b"])


def model_xdeepfm(p, dense, sparse_ids, n_cin=3, n_hidden=3):
    """``XDeepFM`` (MD:121-138) with ``FeatureInput(useLinear=True,
    useAddLinear=True)`` so ``linear_embed`` is ONE ``[B,1,1]`` tensor (the only
    wiring under which ``ScoreLayer(use_add=True)`` at MD:136 is well formed).
    Output is ``sigmoid`` ``[B,1,1]``."""
    sparse, linear = _embed_lists(p, sparse_ids)
    linear = keras_add(linear)                                    # IL:233-234
    cin_in = torch.cat(sparse, dim=1)                             # MD:131
    dnn_in = stack_layer(_dense_list(dense) + sparse)             # MD:132
    cin_out = cin(cin_in, [p[f"cin_w{i}"] for i in range(n_cin)],
                  [p[f"cin_b{i}"] for i in range(n_cin)], p["cin_logit_w"], p["cin_logit_b"])
    dnn_out = dnn_layer(dnn_in, [p[f"dnn_w{i}"] for i in range(n_hidden)],
                        [p[f"dnn_b{i}"] for i in range(n_hidden)], p["dnn_logit_w"], p["dnn_logit_b"])
    return score_layer([linear, cin_out, dnn_out], use_add=True)  # MD:136


def model_nfm(p, dense, sparse_ids, n_hidden=3):
    """``NFM`` (MD:108-119): bi-interaction ``InnerLayer(use_inner=True, use_add=True)`` (MD:112) ->
    ``StackLayer(dense + [cross])`` (MD:113) -> ``DnnLayer(output_dim=1)`` (MD:115) ->
    ``Add(linear_embed + [dnn_fea])`` (MD:116) -> ``ScoreLayer()`` = sigmoid (MD:117).  ``[B,1,1]``."""
    sparse, linear = _embed_lists(p, sparse_ids)
    cross = inner_layer(sparse, use_add=True)                     # [B,1,k]
    dnn_in = stack_layer(_dense_list(dense) + [cross])            # [B, 13 + k]
    dnn_out = dnn_layer(dnn_in, [p[f"dnn_w{i}"] for i in range(n_hidden)],
                        [p[f"dnn_b{i}"] for i in range(n_hidden)], p["dnn_logit_w"], p["dnn_logit_b"])
    return score_layer(keras_add(linear + [dnn_out]))


def model_afm(p, dense, sparse_ids):
    """``AFM`` (MD:141-147): ``InnerLayer()`` list of pairwise products -> ``AttentionBaseLayer()`` ->
    ``ScoreLayer(use_add=True)(linear_embed + [atten_output])`` = sigmoid ``[B,1,1]``."""
    sparse, linear = _embed_lists(p, sparse_ids)
    cross = inner_layer(sparse)                                   # MD:143
    atten = attention_base_layer(cross, p["afm_score_w"], p["afm_score_b"], p["afm_mlp_w"],
                                 p["afm_out_w"], p["afm_out_b"])  # MD:144  [B,1]
    return score_layer(linear + [atten], use_add=True)            # MD:145


def model_pnn(p, dense, sparse_ids, n_hidden=3):
    """``PNN`` (MD:43-56) with ``use_inner=True, use_outer=False`` (``OPnnLayer`` cannot run in the reference:
    ``InnerLayer(use_inner=False)`` reads ``self.dot``, which IL:56 comments out): ``linear_embed`` list +
    ``IPnnLayer()`` = the 325 un-summed pairwise products (IL:68-80) -> ``StackLayer`` (flatten + concat,
    ``[B, F + P*k]``) -> ``DnnLayer`` -> ``Dense(2, softmax)``.  The dense inputs are not used (MD:56)."""
    sparse, linear = _embed_lists(p, sparse_ids)
    cross_fea = list(linear) + inner_layer(sparse)                # MD:48-50
    x = stack_layer(cross_fea)                                    # MD:53
    dnn_ = dnn_layer(x, [p[f"dnn_w{i}"] for i in range(n_hidden)], [p[f"dnn_b{i}"] for i in range(n_hidden)])
    return merge_score_layer(dnn_, p["head_w"], p["head_b"], use_merge=False)


def model_autoint(p, dense, sparse_ids):
    """``AutoInt`` (MD:150-165): one attention block, heads flattened and
    concatenated, ``Dense(2, softmax)``."""
    sparse, _ = _embed_lists(p, sparse_ids)
    x = stack_layer(sparse, use_flat=False, axis=1)               # MD:159
    a = autoint_block(x, p["query_w"], p["key_w"], p["res_w"], p["ln_gamma"], p["ln_beta"])
    heads = [a[h] for h in range(a.shape[0])]                     # MD:162 split+squeeze
    final = stack_layer(heads, use_flat=True, axis=-1)
    return merge_score_layer(final, p["head_w"], p["head_b"], use_merge=False)


def model_autoint_stacked(p, dense, sparse_ids, n_layers: int):
    """EXTENSION with no reference counterpart (the reference wires exactly one attention block, MD:159-163, and a
    second one could not consume its 4-D output, BL:358): ``n_layers`` blocks of ``autoint_block`` stacked the
    standard AutoInt way, re-packing ``[H,B,F,d] -> [B,F,H*d]`` between blocks.  Block 0 uses the reference's
    weight names, block l >= 1 the same names with an ``_l`` suffix.  Each block is BL:356-377 + CL:205-216."""
    sparse, _ = _embed_lists(p, sparse_ids)
    x = stack_layer(sparse, use_flat=False, axis=1)
    a = None
    for l in range(n_layers):
        sfx = "" if l == 0 else f"_{l}"
        a = autoint_block(x, p["query_w" + sfx], p["key_w" + sfx], p["res_w" + sfx], p["ln_gamma" + sfx], p["ln_beta" + sfx])
        if l + 1 < n_layers:
            x = a.permute(1, 2, 0, 3).reshape(a.shape[1], a.shape[2], -1)
    heads = [a[h] for h in range(a.shape[0])]
    final = stack_layer(heads, use_flat=True, axis=-1)
    return merge_score_layer(final, p["head_w"], p["head_b"], use_merge=False)


def glorot_uniform(shape, gen: torch.Generator, dtype=torch.float32):
    """``glorot_uniform``: U(+-sqrt(6/(fan_in+fan_out))) with Keras fan rules
    (2-D: (in,out); >2-D: receptive field * in/out channels).  The TF RNG stream
    is not reproducible here; tests inject explicit weights on both sides."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = (6.0 / (fan_in + fan_out)) ** 0.5
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul(lim).to(dtype)
