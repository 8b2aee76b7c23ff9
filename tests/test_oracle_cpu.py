"""-m "not gpu": the oracle against its fp64 closed forms, torch autograd, and the committed
golden vectors (tests/golden/make_golden.py).  The reference has no tests of its own."""
import itertools
import os

import numpy as np
import torch

from helpers import gen, rel_err
from oracle import kon_oracle as ko

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kon_golden.npz"))


def _t(name):
    return torch.from_numpy(GOLD[name])


def test_golden_embedding_and_grad():
    rows = GOLD["emb_rows"].tolist()
    offs = np.concatenate([[0], np.cumsum(rows)])
    tabs = [_t("emb_tables")[offs[f]:offs[f + 1]] for f in range(3)]
    ids = _t("emb_ids")
    out = torch.cat(ko.sparse_embed([ids[:, f:f + 1] for f in range(3)], tabs, use_flatten=False), 1)
    assert torch.equal(out, _t("emb_out"))
    ur, ug = [], []
    for f in range(3):
        u, g = ko.embedding_grad(GOLD["emb_ids"][:, f], GOLD["emb_dout"][:, f], rows[f])
        ur.append(u + offs[f]); ug.append(g)
    assert np.array_equal(np.concatenate(ur), GOLD["emb_unique_rows"])
    assert np.array_equal(np.concatenate(ug), GOLD["emb_grads"])
    # against autograd of the gather
    w = torch.cat(tabs).clone().requires_grad_(True)
    gl = (ids.long() + torch.tensor(offs[:3])).reshape(-1)
    (w[gl].reshape(-1, 3, 8) * _t("emb_dout")).sum().backward()
    dense = np.zeros_like(GOLD["emb_tables"]); dense[GOLD["emb_unique_rows"]] = GOLD["emb_grads"]
    assert rel_err(w.grad, torch.from_numpy(dense)) < 1e-6


def test_golden_fm_and_closed_form():
    rows = GOLD["emb_rows"].tolist()
    offs = np.concatenate([[0], np.cumsum(rows)])
    ids = _t("emb_ids")
    idl = [ids[:, f:f + 1] for f in range(3)]
    emb = ko.sparse_embed(idl, [_t("emb_tables")[offs[f]:offs[f + 1]] for f in range(3)], use_flatten=False)
    lin = ko.sparse_embed(idl, [_t("emb_lins")[offs[f]:offs[f + 1]] for f in range(3)], use_flatten=False)
    y = ko.fm_layer(emb, lin)
    assert y.shape == (33, 1, 8)                       # the "FM" keeps the embedding axis (IL:161-170)
    assert torch.equal(y[:, 0], _t("fm_out"))
    assert rel_err(y[:, 0], _t("fm_out_f64")) < 1e-5


def test_fm_pair_order_and_count():
    g = gen(3)
    xs = [torch.randn(4, 1, 3, generator=g, dtype=torch.float64) for _ in range(26)]
    pairs = ko.inner_layer(xs)
    assert len(pairs) == 325
    assert torch.equal(pairs[1], xs[0] * xs[2])         # itertools.combinations order
    v = torch.cat(xs, 1)
    ref = ko.fm_closed_form(v, torch.zeros(4, 26, dtype=torch.float64))
    assert rel_err(ko.inner_layer(xs, use_add=True)[:, 0], ref) < 1e-12


def test_golden_cross():
    x, w, b = _t("cross_x"), _t("cross_w"), _t("cross_b")
    out = ko.cross_layer(x, list(w), list(b))
    assert out.shape == (33, 29, 1)
    assert torch.equal(out, _t("cross_out"))
    # closed form in fp64
    x0 = x.double(); xl = x0
    for l in range(3):
        s = (xl * w[l, :, 0].double()).sum(1, keepdim=True)
        xl = x0 * s + xl + b[l, :, 0].double()
    assert rel_err(out[..., 0], xl) < 1e-5


def test_golden_cin_mirror_equals_closed_form():
    x0 = _t("cin_x0")
    ws, bs = [_t("cin_w0"), _t("cin_w1")], [_t("cin_b0"), _t("cin_b1")]
    pooled = ko.cin(x0, ws, bs, return_pooled=True)
    assert pooled.shape == (33, 2 * 4)                 # pools over the feature maps -> [B,D] per layer
    assert torch.equal(pooled, _t("cin_pooled"))
    cf, zs = ko.cin_closed_form(x0.double(), [w.double() for w in ws], [b.double() for b in bs])
    assert rel_err(cf, _t("cin_pooled_f64")) < 1e-12
    assert rel_err(pooled, cf) < 1e-5
    # channel index is h*m + i (IL:317-318): perturb one weight and watch the right product move
    m, h, i, o = 5, 2, 3, 1
    w2 = ws[0].double().clone(); w2[0, h * m + i, o] += 1.0
    _, z2 = ko.cin_closed_form(x0.double(), [w2, ws[1].double()], [b.double() for b in bs])
    dz = z2[0] - zs[0]
    expect = x0[:, h].double() * x0[:, i].double()      # layer 1: pre == x0
    assert rel_err(dz[:, :, o], expect) < 1e-12


def test_golden_attention_block_semantics():
    x, wq, wk, wr = _t("attn_x"), _t("attn_wq"), _t("attn_wk"), _t("attn_wr")
    gam, bet = _t("attn_gamma"), _t("attn_beta")
    y = ko.autoint_block(x, wq, wk, wr, gam, bet)
    assert y.shape == (2, 33, 6, 4)
    assert torch.equal(y, _t("attn_out"))
    # sigmoid (not softmax), V = X key_w, LN eps 1e-3 before the residual, ReLU last
    xd = x.double()
    for h in range(2):
        q, k = xd @ wq[:, h].double(), xd @ wk[:, h].double()
        s = torch.sigmoid(q @ k.transpose(1, 2) / 2.0)
        o = s @ k
        mu, var = o.mean(-1, keepdim=True), o.var(-1, unbiased=False, keepdim=True)
        ln = (o - mu) / torch.sqrt(var + 1e-3) * gam.double() + bet.double()
        ref = torch.relu(ln + xd @ wr[:, h].double())
        assert rel_err(y[h], ref) < 1e-5


def test_keras_add_rank_expansion():
    a, b = torch.ones(5, 1, 1), torch.ones(5, 1)
    assert ko.keras_add([a, b]).shape == (5, 1, 1)


def test_oracle_models_run_and_shapes():
    g = gen(7)
    rows = [5, 9, 4]
    B, k = 12, 4
    p = {}
    for f, r in enumerate(rows):
        p[f"emb_{f}"] = torch.randn(r, k, generator=g); p[f"lin_{f}"] = torch.randn(r, 1, generator=g)
    D = 13 + 3 * k
    dims = [D, 8, 6, 5]
    for i in range(3):
        p[f"dnn_w{i}"] = torch.randn(dims[i], dims[i + 1], generator=g) * 0.2
        p[f"dnn_b{i}"] = torch.randn(dims[i + 1], generator=g) * 0.1
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1)
    dense = torch.rand(B, 13, generator=g)
    p["head_w"], p["head_b"] = torch.randn(k + 5, 2, generator=g), torch.zeros(2)
    assert ko.model_deepfm(p, dense, ids).shape == (B, 2)
    hp = 3
    for i, n in enumerate((4, 3, 2)):
        p[f"cin_w{i}"] = torch.randn(1, hp * 3, n, generator=g) * 0.3; p[f"cin_b{i}"] = torch.zeros(n); hp = n
    p["cin_logit_w"], p["cin_logit_b"] = torch.randn(3 * k, 1, generator=g), torch.zeros(1)
    p["dnn_logit_w"], p["dnn_logit_b"] = torch.randn(5, 1, generator=g), torch.zeros(1)
    out = ko.model_xdeepfm(p, dense, ids)
    assert out.shape == (B, 1, 1) and float(out.min()) > 0 and float(out.max()) < 1


def test_nfm_op_mirror_equals_closed_form():
    """NFM (MD:108-119): the 325-product bi-interaction ``InnerLayer(use_add=True)`` equals
    ``0.5((sum v)^2 - sum v^2)``; the rest is Dense/ReLU and a Keras Add with rank expansion."""
    g = gen(31)
    rows, k, B = [7, 30, 5, 11], 6, 19
    p = {}
    for f, r in enumerate(rows):
        p[f"emb_{f}"] = torch.randn(r, k, generator=g, dtype=torch.float64)
        p[f"lin_{f}"] = torch.randn(r, 1, generator=g, dtype=torch.float64)
    dims = [13 + k, 9, 7, 5]
    for i in range(3):
        p[f"dnn_w{i}"] = torch.randn(dims[i], dims[i + 1], generator=g, dtype=torch.float64) * 0.3
        p[f"dnn_b{i}"] = torch.randn(dims[i + 1], generator=g, dtype=torch.float64) * 0.1
    p["dnn_logit_w"] = torch.randn(5, 1, generator=g, dtype=torch.float64)
    p["dnn_logit_b"] = torch.randn(1, generator=g, dtype=torch.float64)
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1)
    dense = torch.rand(B, 13, generator=g, dtype=torch.float64)
    out = ko.model_nfm(p, dense, ids)
    assert out.shape == (B, 1, 1)
    v = torch.stack([p[f"emb_{f}"][ids[:, f]] for f in range(len(rows))], 1)           # [B,F,k]
    bi = 0.5 * (v.sum(1) ** 2 - (v * v).sum(1))
    x = torch.cat([dense, bi], 1)
    for i in range(3):
        x = torch.relu(x @ p[f"dnn_w{i}"] + p[f"dnn_b{i}"])
    logit = x @ p["dnn_logit_w"] + p["dnn_logit_b"]
    lin = sum(p[f"lin_{f}"][ids[:, f]] for f in range(len(rows)))                       # [B,1]
    ref = torch.sigmoid(lin + logit).view(B, 1, 1)
    assert rel_err(out, ref) < 1e-12


def test_golden_head_and_nfm():
    """Frozen vectors of the 2-unit head (MergeScoreLayer, CL:86-100) and of NFM end to end (MD:108-119)."""
    x1, x2, w, b = _t("head_x1"), _t("head_x2"), _t("head_w"), _t("head_b")
    logits = ko.keras_dense(torch.cat([x1, x2], 1), w, b)
    assert torch.equal(logits, _t("head_logits"))
    assert torch.equal(ko.merge_score_layer([x1, x2], w, b), _t("head_softmax"))
    assert rel_err(torch.softmax(logits.double(), -1), _t("head_softmax").double()) < 1e-6
    rows = GOLD["emb_rows"].tolist()
    offs = np.concatenate([[0], np.cumsum(rows)])
    p = {}
    for f in range(3):
        p[f"emb_{f}"] = _t("emb_tables")[offs[f]:offs[f + 1]]
        p[f"lin_{f}"] = _t("emb_lins")[offs[f]:offs[f + 1]]
    for i in range(3):
        p[f"dnn_w{i}"], p[f"dnn_b{i}"] = _t(f"nfm_dnn_w{i}"), _t(f"nfm_dnn_b{i}")
    p["dnn_logit_w"], p["dnn_logit_b"] = _t("nfm_logit_w"), torch.zeros(1)
    out = ko.model_nfm(p, _t("nfm_dense"), _t("emb_ids"))
    assert torch.equal(out, _t("nfm_out"))


# ------------------------------------------------------------------ plain-C restatement (oracle/kon_oracle_c.c)
def test_c_oracle_agrees_with_torch_oracle_and_golden():
    """An independent plain-C restatement of the byte-exact parts (gather, gradient routing + segment sums in
    sample order) and of the fp32 op order of FmLayer / CrossLayer must reproduce the torch oracle's bits
    (gather, routing, FM) resp. its values to fp32 rounding (cross: the MatMul's summation order is a BLAS detail),
    and the committed golden vectors."""
    from oracle import c_oracle as co
    rows = GOLD["emb_rows"].tolist()
    offs = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    assert np.array_equal(co.embed_gather(GOLD["emb_tables"], offs, GOLD["emb_ids"]), GOLD["emb_out"])
    ur, ug = co.embed_grad(offs, GOLD["emb_ids"], GOLD["emb_dout"])
    assert np.array_equal(ur, GOLD["emb_unique_rows"]) and np.array_equal(ug, GOLD["emb_grads"])
    lin = np.stack([GOLD["emb_lins"][offs[f] + GOLD["emb_ids"][:, f], 0] for f in range(3)], 1)
    assert np.array_equal(co.fm(GOLD["emb_out"], lin), GOLD["fm_out"])              # reference op order, bit for bit
    got = co.cross(GOLD["cross_x"], GOLD["cross_w"][..., 0], GOLD["cross_b"][..., 0])
    assert rel_err(torch.from_numpy(got), _t("cross_out")[..., 0]) < 2e-6
    # a larger random case incl. heavy duplicates and the 26-field / 325-pair FM
    g = gen(5)
    rows = [3, 40, 1, 1000, 7] + [11] * 21
    offs = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    B, k = 301, 16
    tabs = [torch.randn(r, k, generator=g) for r in rows]
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
    emb = ko.sparse_embed([ids[:, f:f + 1] for f in range(26)], tabs, use_flatten=False)
    v = torch.cat(emb, 1)
    assert np.array_equal(co.embed_gather(torch.cat(tabs).numpy(), offs, ids.numpy()), v.numpy())
    lins = [torch.randn(B, 1, 1, generator=g) for _ in range(26)]
    ref = ko.fm_layer(emb, lins)[:, 0]
    assert np.array_equal(co.fm(v.numpy(), torch.cat(lins, 1)[..., 0].numpy()), ref.numpy())
    d_out = torch.randn(B, 26, k, generator=g)
    ur, ug = co.embed_grad(offs, ids.numpy(), d_out.numpy())
    eu, eg = [], []
    for f in range(26):
        u, gr = ko.embedding_grad(ids[:, f].numpy(), d_out[:, f].numpy(), rows[f])
        eu.append(u + offs[f]); eg.append(gr)
    assert np.array_equal(ur, np.concatenate(eu)) and np.array_equal(ug, np.concatenate(eg))


def test_c_oracle_cin_and_attention_against_golden():
    """The C restatement of CIN.call (IL:310-327) and of the AutoInt block (BL:292-311, 356-377, CL:201-226)
    against the committed vectors (fp32 op-mirror and fp64 closed form of the torch oracle)."""
    from oracle import c_oracle as co
    ws = [GOLD[f"cin_w{i}"][0] for i in range(2)]
    bs = [GOLD[f"cin_b{i}"] for i in range(2)]
    pooled = co.cin(GOLD["cin_x0"], ws, bs)
    assert rel_err(torch.from_numpy(pooled), _t("cin_pooled_f64")) < 2e-6
    assert rel_err(torch.from_numpy(pooled), _t("cin_pooled")) < 1e-5
    y = co.autoint_block(GOLD["attn_x"], GOLD["attn_wq"], GOLD["attn_wk"], GOLD["attn_wr"], GOLD["attn_gamma"],
                         GOLD["attn_beta"])
    assert y.shape == GOLD["attn_out"].shape
    assert rel_err(torch.from_numpy(y), _t("attn_out")) < 1e-5
