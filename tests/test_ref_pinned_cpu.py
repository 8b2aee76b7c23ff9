"""The oracle, pinned to the reference's own source.

tests/golden/ref_layers.npz / ref_models.npz are outputs of the UNMODIFIED reference classes
(/root/reference/kon/...) executed over the eager tensorflow shim (oracle/_ref_shim); see
tests/golden/make_ref_golden.py.  These tests assert that the restatement in oracle/kon_oracle.py
reproduces them -- BIT FOR BIT in fp32 (both sides are torch-CPU fp32 in the reference's op
order, so any difference is a difference in what is computed, not in rounding) -- and, where
/root/reference is present, that the committed fixtures are exactly what the reference
produces today.
"""
import os

import numpy as np
import pytest
import torch

from helpers import cin26_weights, ref_case
from oracle import kon_oracle as ko
from oracle import run_reference


def eq(a, b, what=""):
    a, b = a.detach(), b.detach()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(a, b), f"{what}: max |diff| {(a.double() - b.double()).abs().max().item():.3e}"


@pytest.mark.skipif(not run_reference.available(), reason="reference tree not present (GPU box)")
def test_fixtures_are_what_the_reference_produces():
    import importlib.util
    p = os.path.join(os.path.dirname(__file__), "golden", "make_ref_golden.py")
    spec = importlib.util.spec_from_file_location("make_ref_golden", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(verbose=False) == []


@pytest.mark.skipif(not run_reference.available(), reason="reference tree not present (GPU box)")
def test_reference_modules_come_from_the_reference_tree():
    ns = run_reference.load()
    for mod in (ns.IL, ns.CL, ns.BL, ns.MD, ns.DP):
        assert mod.__file__.startswith(run_reference.REFERENCE_ROOT)
    assert ns.tf.__version__.endswith("-shim")


# ---- a1-a3 ---------------------------------------------------------------------------------------
def test_sparse_embed_bit_exact():
    c = ref_case("layers", "sparse_embed")
    F = c["in/ids"].shape[1]
    ids = [c["in/ids"][:, f:f + 1] for f in range(F)]                    # float32 ids, as in the reference
    tabs = [c.w["emb_%d" % f] for f in range(F)]
    lins = [c.w["lin_%d" % f] for f in range(F)]
    eq(torch.stack(ko.sparse_embed(ids, tabs, use_flatten=True), 1), c["out/flat"], "flat")
    eq(torch.stack(ko.sparse_embed(ids, tabs, use_flatten=False), 1), c["out/noflat"], "noflat")
    eq(ko.sparse_embed(ids, tabs, use_flatten=False, use_add=True), c["out/add"], "add")
    eq(torch.stack(ko.sparse_embed(ids, lins, use_flatten=False), 1), c["out/lin"], "lin")
    eq(ko.sparse_embed(ids, lins, use_flatten=False, use_add=True), c["out/linadd"], "linadd")


def test_seq_embed_and_sum_pool_bit_exact():
    c = ref_case("layers", "seq_embed")
    ids = [c["in/ids"][:, f] for f in range(2)]
    tabs = [c.w["emb_%d" % f] for f in range(2)]
    emb = ko.sparse_embed(ids, tabs, use_flatten=False)
    eq(torch.stack(emb, 1), c["out/emb"], "emb")
    eq(torch.stack(ko.seq_base_layer(emb), 1), c["out/pooled"], "pooled")
    eq(torch.stack([i != 0 for i in ids], 1).to(torch.uint8), c["out/mask"], "mask")


# ---- a5-a6 ---------------------------------------------------------------------------------------
def test_inner_and_fm_bit_exact():
    c = ref_case("layers", "fm")
    v, lin = c["in/v"], c["in/lin"]
    F = v.shape[1]
    vl = [v[:, f:f + 1].clone().requires_grad_(True) for f in range(F)]
    ll = [lin[:, f].reshape(-1, 1, 1).clone().requires_grad_(True) for f in range(F)]
    eq(torch.stack(ko.inner_layer(vl), 1)[:, :, 0], c["out/pairs"], "pairs")
    eq(ko.inner_layer(vl, use_add=True), c["out/inner_add"], "inner_add")
    fm = ko.fm_layer(vl, ll)
    eq(fm, c["out/fm"], "fm")
    (fm * c["in/gy"]).sum().backward()
    eq(torch.stack([t.grad for t in vl], 1)[:, :, 0], c["grad/v"], "dv")
    eq(torch.stack([t.grad for t in ll], 1)[:, :, 0, 0], c["grad/lin"], "dlin")
    # closed form vs the reference in fp64
    cf = ko.fm_closed_form(v.double(), lin.double())
    assert (cf - c["out64/fm"][:, 0]).abs().max() < 1e-12


def test_fm_26_fields_bit_exact():
    c = ref_case("layers", "fm26")
    v, lin = c["in/v"], c["in/lin"]
    vl = [v[:, f:f + 1] for f in range(26)]
    ll = [lin[:, f].reshape(-1, 1, 1) for f in range(26)]
    eq(ko.fm_layer(vl, ll), c["out/fm"], "fm26")


# ---- a7 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cross", "cross6"])
def test_cross_bit_exact(name):
    c = ref_case("layers", name)
    L = len([k for k in c.w if k.startswith("outer_weight_")])
    x = c["in/x"].clone().requires_grad_(True)
    ws = [c.w["outer_weight_%d" % i].clone().requires_grad_(True) for i in range(L)]
    bs = [c.w["outer_bias_%d" % i].clone().requires_grad_(True) for i in range(L)]
    y = ko.cross_layer(x, ws, bs)
    eq(y, c["out/y"], "y")
    (y * c["in/gy"]).sum().backward()
    eq(x.grad, c["grad/x"], "dx")
    for i in range(L):
        eq(ws[i].grad, c["grad/outer_weight_%d" % i], "dw%d" % i)
        eq(bs[i].grad, c["grad/outer_bias_%d" % i], "db%d" % i)


# ---- a8 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cin", "cin_pooled"])
def test_cin_bit_exact(name):
    c = ref_case("layers", name)
    n = len(c["meta/hs"])
    x0 = c["in/x0"].clone().requires_grad_(True)
    ws = [c.w["cin_w%d" % i].clone().requires_grad_(True) for i in range(n)]
    bs = [c.w["cin_b%d" % i].clone().requires_grad_(True) for i in range(n)]
    if name == "cin":
        lw = c.w["cin_logit_w"].clone().requires_grad_(True)
        y = ko.cin(x0, ws, bs, lw, c.w["cin_logit_b"])
    else:
        y = ko.cin(x0, ws, bs, return_pooled=True)
    eq(y, c["out/y"], "y")
    (y * c["in/gy"]).sum().backward()
    eq(x0.grad, c["grad/x0"], "dx0")
    for i in range(n):
        eq(ws[i].grad, c["grad/cin_w%d" % i], "dw%d" % i)
        eq(bs[i].grad, c["grad/cin_b%d" % i], "db%d" % i)
    # the closed form the kernels implement == the reference, in fp64
    pooled64, _ = ko.cin_closed_form(c["in/x0"].double(), [w.detach().double() for w in ws], [b.detach().double() for b in bs])
    if name == "cin":
        pooled64 = pooled64 @ c.w["cin_logit_w"].double() + c.w["cin_logit_b"].double()
    assert (pooled64 - c["out64/y"]).abs().max() < 1e-10


def test_cin_26_fields_200_maps_bit_exact():
    c = ref_case("layers", "cin26")
    ws, bs, lw, lb = cin26_weights(int(c["meta/seed"]))
    x0 = c["in/x0"].clone().requires_grad_(True)
    ws = [w.requires_grad_(True) for w in ws]
    bs = [b.requires_grad_(True) for b in bs]
    eq(lw, c.w["cin_logit_w"], "seeded logit weights")
    y = ko.cin(x0, ws, bs, lw, lb)
    eq(y, c["out/y"], "y")
    (y * c["in/gy"]).sum().backward()
    eq(x0.grad, c["grad/x0"], "dx0")
    for i in range(3):
        eq(ws[i].grad[:, :64], c["grad/cin_w%d_head" % i], "dw%d head" % i)
        eq(ws[i].grad.sum(dim=1), c["grad/cin_w%d_colsum" % i], "dw%d colsum" % i)
        eq(bs[i].grad, c["grad/cin_b%d" % i], "db%d" % i)


# ---- a9 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,kw", [("plain", {}), ("scale", dict(use_scale=True)),
                                    ("mask1", dict(use_scale=True, mask_mod=1)),
                                    ("mask2", dict(use_scale=True, mask_mod=2))])
def test_product_attention_bit_exact(tag, kw):
    c = ref_case("layers", "product_attention")
    q, k, v = (c["in/" + n].clone().requires_grad_(True) for n in "qkv")
    mask = None
    if tag == "mask1":
        mask = c["in/mask1"]
    if tag == "mask2":
        mask = c["in/mask2"].bool()
    y = ko.product_attention(q, k, v, mask=mask, **kw)
    eq(y, c["out/" + tag], tag)
    (y * c["in/gy"]).sum().backward()
    eq(q.grad, c["grad/%s_q" % tag], "dq")
    eq(k.grad, c["grad/%s_k" % tag], "dk")
    eq(v.grad, c["grad/%s_v" % tag], "dv")


# ---- a10 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mha", "mha_h3", "mha_mask2"])
def test_mult_head_attention_and_block_bit_exact(name):
    c = ref_case("layers", name)
    w = {k: v.clone().requires_grad_(True) for k, v in c.w.items()}
    mask = c["in/mask"].bool() if "in/mask" in c else None
    mod = 2 if mask is not None else 1
    x = c["in/x"].clone().requires_grad_(True)
    atten_v, res = ko.mult_head_attention(x, w["query_w"], w["key_w"], w["res_w"], w["ln_gamma"], w["ln_beta"],
                                          mask=mask, atten_mask_mod=mod)
    eq(atten_v, c["out/atten_v"], "atten_v")
    eq(res, c["out/res"], "res")
    # DnnLayer.call (CL:201-226) never hands a mask to its hidden layer: the wrapped block is unmasked
    blk = ko.autoint_block(x, w["query_w"], w["key_w"], w["res_w"], w["ln_gamma"], w["ln_beta"])
    eq(blk, c["out/block"], "block")
    (blk * c["in/gy"]).sum().backward()
    eq(x.grad, c["grad/x"], "dx")
    for n in ("query_w", "key_w", "res_w", "ln_gamma", "ln_beta"):
        eq(w[n].grad, c["grad/" + n], "d" + n)


def test_mult_head_attention_one_head_is_squeezed():
    c = ref_case("layers", "mha_h1")
    w = c.w
    out = ko.mult_head_attention(c["in/x"], w["query_w"], w["key_w"], w["res_w"], w["ln_gamma"], w["ln_beta"])
    assert out.dim() == 3                                               # BL:374-375
    eq(out, c["out/atten_v"], "atten_v")


# ---- a12 -----------------------------------------------------------------------------------------
def test_dnn_layer_with_firing_residual_bit_exact():
    c = ref_case("layers", "dnn")
    n = len([k for k in c.w if k.startswith("dnn_w")])
    x = c["in/x"].clone().requires_grad_(True)
    ws = [c.w["dnn_w%d" % i].clone().requires_grad_(True) for i in range(n)]
    y = ko.dnn_layer(x, ws, [c.w["dnn_b%d" % i] for i in range(n)], c.w["dnn_logit_w"], c.w["dnn_logit_b"])
    eq(y, c["out/y"], "y")
    y.sum().backward()
    eq(x.grad, c["grad/x"], "dx")
    for i in range(n):
        eq(ws[i].grad, c["grad/dnn_w%d" % i], "dw%d" % i)


def test_heads_bit_exact():
    c = ref_case("layers", "heads")
    eq(ko.merge_score_layer([c["in/x1"], c["in/x2"]], c.w["head_w"], c.w["head_b"]), c["out/merge"], "merge")
    eq(ko.score_layer([c["in/s1"], c["in/s2"], c["in/s3"]], use_add=True), c["out/score_add"], "score_add")
    eq(ko.score_layer(c["in/s1"]), c["out/score"], "score")


# ---- the builders, end to end: forward, loss, every gradient -----------------------------------------
MODELS = {
    "fm": (ko.model_fm, {}),
    "deepfm": (ko.model_deepfm, {}),
    "dcn": (ko.model_dcn, dict(cross_hidden=6)),
    "xdeepfm": (ko.model_xdeepfm, {}),
    "autoint": (ko.model_autoint, {}),
    "nfm": (ko.model_nfm, {}),
    "afm": (ko.model_afm, {}),
    "pnn": (ko.model_pnn, {}),
}


@pytest.mark.parametrize("name", sorted(MODELS))
def test_model_builders_bit_exact(name):
    c = ref_case("models", name)
    fn, kw = MODELS[name]
    p = ko.OracleParams({k: v.clone().requires_grad_(True) for k, v in c.w.items()})
    out = fn(p, c["in/dense"], c["in/ids"], **kw)
    eq(out, c["out/y"], "out")
    loss = ko.binary_crossentropy(c["in/labels"], out)
    eq(loss, c["out/loss"], "loss")
    loss.backward()
    g = c.grads()
    assert g, "fixture has no gradients"
    for n, ref in g.items():
        assert p[n].grad is not None, f"oracle produced no gradient for {n}"
        eq(p[n].grad, ref, "d" + n)
    # weights the reference never reads get no gradient on either side (value_w, BL:360)
    for n in p:
        if n not in g:
            assert p[n].grad is None or float(p[n].grad.abs().max()) == 0.0, n
    # fp32 reference vs its own fp64 run: the error floor every fp32 tolerance is judged against
    assert (out.detach().double() - c["out64/y"]).abs().max() < 1e-5


def test_l2_regulariser_of_the_tables_is_what_keras_adds():
    """IL:217: l2(emb_reg) on every cross-embedding table -> sum_f emb_reg * sum(T_f^2)."""
    c = ref_case("models", "deepfm")
    want = float(sum(1e-8 * (v.double() ** 2).sum() for k, v in c.w.items() if k.startswith("emb_")))
    assert abs(float(c["out/reg_loss"]) - want) < 1e-6 * want
    assert want > 0
