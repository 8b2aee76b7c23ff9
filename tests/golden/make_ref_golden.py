"""Generates tests/golden/ref_layers.npz and tests/golden/ref_models.npz by EXECUTING THE
UNMODIFIED REFERENCE CLASSES from /root/reference (interactive_layer.py, core_layer.py,
behavior_layer.py, models.py, data_prepare.py) over the eager `tensorflow` shim in
oracle/_ref_shim (TensorFlow itself is not installable here; oracle/_ref_shim/README.md lists
exactly which Keras semantics the shim supplies).

    python tests/golden/make_ref_golden.py            # rewrite the fixtures
    python tests/golden/make_ref_golden.py --check    # regenerate in memory, compare with the files

Every case stores: the inputs fed (`<case>/in/...`), every weight the reference layer created
(`<case>/w/<oracle name>`), the fp32 outputs (`<case>/out...`), the same outputs from an fp64
run of the same reference code (`<case>/out64...`, the "truth" tolerances are judged against)
and -- where the case has a loss -- the gradients of that loss w.r.t. every weight and input
(`<case>/grad/...`, torch autograd over the reference's own op sequence standing in for
tf.GradientTape).  Weight names are the ones oracle/kon_oracle.py and
`load_reference_params` use, so one loader serves the oracle and the CUDA tests.

The GPU box has no /root/reference: the -m gpu tests only read the committed .npz files.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import run_reference  # noqa: E402

LAYERS_NPZ = os.path.join(HERE, "ref_layers.npz")
MODELS_NPZ = os.path.join(HERE, "ref_models.npz")

N_DENSE = 13


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().as_subclass(torch.Tensor).numpy().copy()
    return np.asarray(t)


class Ctx:
    """One run of the reference at a given float width."""

    def __init__(self, ns, f64: bool):
        self.ns, self.tf, self.f64 = ns, ns.tf, f64
        self.dt = torch.float64 if f64 else torch.float32

    def __enter__(self):
        self.tf._FLOATX[0] = self.dt
        self.tf.keras.reset_layers()
        self.tf.keras.initializers.reseed(2020)
        self.tf.keras.initializers.set_hook(None)
        return self

    def __exit__(self, *a):
        self.tf._FLOATX[0] = torch.float32
        self.tf.keras.initializers.set_hook(None)

    def t(self, a, grad=False):
        x = torch.as_tensor(np.asarray(a))
        if x.dtype in (torch.float32, torch.float64):
            x = x.to(self.dt)
        return x.clone().requires_grad_(True) if grad else x


def both(ns, fn):
    """Run `fn(ctx) -> dict` in fp32 and fp64; fp64 results of keys starting with 'out' are
    stored as 'out64...' (the "truth" that tolerances are judged against); everything else,
    gradients included, comes from the fp32 run -- the reference's own arithmetic."""
    with Ctx(ns, False) as c:
        r32 = fn(c)
    with Ctx(ns, True) as c:
        r64 = fn(c)
    out = {k: _np(v) for k, v in r32.items()}
    for k, v in r64.items():
        if k.startswith("out"):
            out["out64" + k[3:]] = _np(v)
    return out


def set_weights(weights, arrays):
    with torch.no_grad():
        for w, a in zip(weights, arrays):
            w.copy_(torch.as_tensor(np.asarray(a)).to(w.dtype))


# ------------------------------------------------------------------------------------------------
# layer cases
# ------------------------------------------------------------------------------------------------
def layer_cases(ns):
    rs = np.random.RandomState(20201)
    cases = {}
    IL, CL, BL, DP = ns.IL, ns.CL, ns.BL, ns.DP
    dp = DP.data_prepare()

    def sfea(name, rows, k, L=1, mask_zero=False):
        return dp.sparseFea(fea_name=name, word_size=rows, input_dim=rows, cross_unit=k, linear_unit=1,
                            pre_weight=None, mask_zero=mask_zero, is_trainable=True, input_length=L,
                            sample_num=None, batch_size=None, emb_reg=1e-8)

    # ---- a1/a2: SparseEmbed (IL:196-247) ------------------------------------------------------
    B, k = 37, 8
    rows = [7, 50, 3, 19, 1]
    ids = np.stack([rs.randint(0, r, B) for r in rows], 1).astype(np.float32)       # float32 ids, DP:290-292
    tabs = [rs.randn(r, k).astype(np.float32) for r in rows]
    lins = [rs.randn(r, 1).astype(np.float32) for r in rows]
    info = [sfea("s%d" % f, r, k) for f, r in enumerate(rows)]

    def embed_case(c):
        idl = [c.t(ids[:, f:f + 1]) for f in range(len(rows))]
        res = {"in/ids": ids}
        for tag, kw, ws in (("flat", dict(use_flatten=True), tabs),
                            ("noflat", dict(use_flatten=False), tabs),
                            ("add", dict(use_flatten=False, use_add=True), tabs),
                            ("lin", dict(is_linear=True, use_flatten=False), lins),
                            ("linadd", dict(is_linear=True, use_flatten=False, use_add=True), lins)):
            layer = IL.SparseEmbed(info, **kw)
            layer(idl)                                        # builds the 5 Keras Embedding layers
            for e, w in zip(layer.embed, ws):
                e.set_weights([w])
            out = layer(idl)
            res["out/" + tag] = out if isinstance(out, torch.Tensor) else torch.stack(list(out), 1)
        for f in range(len(rows)):
            res["w/emb_%d" % f] = tabs[f]
            res["w/lin_%d" % f] = lins[f]
        return res
    cases["sparse_embed"] = both(ns, embed_case)

    # sequence features: input_length L, mask_zero (DP:74) + SeqBaseLayer sum pooling (BL:45-46)
    Ls, rows_s = 6, [11, 4]
    ids_s = np.stack([rs.randint(0, r, (B, Ls)) for r in rows_s], 1).astype(np.float32)   # [B,2,L]
    ids_s[:, :, -2:] = 0                                                                  # padding id 0
    tabs_s = [rs.randn(r, k).astype(np.float32) for r in rows_s]
    info_s = [sfea("q%d" % f, r, k, L=Ls, mask_zero=True) for f, r in enumerate(rows_s)]

    def seq_case(c):
        idl = [c.t(ids_s[:, f]) for f in range(2)]
        layer = IL.SparseEmbed(info_s, support_masking=True, mask_zero=True, is_linear=False, use_flatten=False)
        layer(idl)
        for e, w in zip(layer.embed, tabs_s):
            e.set_weights([w])
        emb, masks = layer(idl)
        pooled = BL.SeqBaseLayer()(emb)
        res = {"in/ids": ids_s, "out/emb": torch.stack(list(emb), 1),
               "out/mask": torch.stack(list(masks), 1).to(torch.uint8),
               "out/pooled": torch.stack(list(pooled), 1)}
        for f in range(2):
            res["w/emb_%d" % f] = tabs_s[f]
        return res
    cases["seq_embed"] = both(ns, seq_case)

    # ---- a5/a6: InnerLayer (IL:59-66), FmLayer (IL:161-170) ------------------------------------
    F = 6
    v = rs.randn(B, F, 1, k).astype(np.float32)
    lin = rs.randn(B, F, 1, 1).astype(np.float32)
    gy = rs.randn(B, 1, k).astype(np.float32)

    def fm_case(c):
        vl = [c.t(v[:, f], grad=True) for f in range(F)]
        ll = [c.t(lin[:, f], grad=True) for f in range(F)]
        pairs = IL.InnerLayer()(vl)                           # list of F(F-1)/2 [B,1,k]
        summed = IL.InnerLayer(use_inner=True, use_add=True)(vl)
        fm = IL.FmLayer()([vl, ll])
        (fm * c.t(gy)).sum().backward()
        return {"in/v": v[:, :, 0], "in/lin": lin[:, :, 0, 0], "in/gy": gy,
                "out/pairs": torch.stack(list(pairs), 1)[:, :, 0], "out/inner_add": summed, "out/fm": fm,
                "grad/v": torch.stack([t.grad for t in vl], 1)[:, :, 0],
                "grad/lin": torch.stack([t.grad for t in ll], 1)[:, :, 0, 0]}
    cases["fm"] = both(ns, fm_case)

    # 26-field FM (the 325-pair rounding order at the configuration the benches use)
    v26 = rs.randn(B, 26, 1, 16).astype(np.float32)
    l26 = rs.randn(B, 26, 1, 1).astype(np.float32)

    def fm26_case(c):
        vl = [c.t(v26[:, f]) for f in range(26)]
        ll = [c.t(l26[:, f]) for f in range(26)]
        return {"in/v": v26[:, :, 0], "in/lin": l26[:, :, 0, 0], "out/fm": IL.FmLayer()([vl, ll])}
    cases["fm26"] = both(ns, fm26_case)

    # ---- a7: CrossLayer (IL:264-282) -------------------------------------------------------------
    for tag, (D, Lc) in (("cross", (29, 3)), ("cross6", (13 + 26 * 4, 6))):
        x = rs.randn(B, D).astype(np.float32)
        cw = [(rs.randn(D, 1) * 0.3).astype(np.float32) for _ in range(Lc)]
        cb = [(rs.randn(D, 1) * 0.1).astype(np.float32) for _ in range(Lc)]
        g = rs.randn(B, D, 1).astype(np.float32)

        def cross_case(c, x=x, cw=cw, cb=cb, g=g, Lc=Lc):
            layer = IL.CrossLayer(cross_hidden=Lc)
            xt = c.t(x, grad=True)
            layer(xt)
            set_weights(layer.kernel, cw)
            set_weights(layer.bias, cb)
            y = layer(xt)
            (y * c.t(g)).sum().backward()
            res = {"in/x": x, "in/gy": g, "out/y": y, "grad/x": xt.grad}
            for i in range(Lc):
                res["w/outer_weight_%d" % i], res["w/outer_bias_%d" % i] = cw[i], cb[i]
                res["grad/outer_weight_%d" % i], res["grad/outer_bias_%d" % i] = layer.kernel[i].grad, layer.bias[i].grad
            return res
        cases[tag] = both(ns, cross_case)

    # ---- a8: CIN (IL:296-327) --------------------------------------------------------------------
    def cin_weights(m, hs, D, seed, scale=1.0):
        r = np.random.RandomState(seed)
        ws, bs, hp = [], [], m
        for n in hs:
            lim = scale * (6.0 / (hp * m + n)) ** 0.5
            ws.append(r.uniform(-lim, lim, (1, hp * m, n)).astype(np.float32))
            bs.append((r.randn(n) * 0.05).astype(np.float32))
            hp = n
        lw = (r.randn(len(hs) * D, 1) * 0.2).astype(np.float32)
        lb = (r.randn(1) * 0.1).astype(np.float32)
        return ws, bs, lw, lb

    def make_cin_case(Bc, m, D, hs, seed, store_w, out_dim=1):
        x0 = np.random.RandomState(seed + 1).randn(Bc, m, D).astype(np.float32)
        ws, bs, lw, lb = cin_weights(m, hs, D, seed)
        g = np.random.RandomState(seed + 2).randn(Bc, 1 if out_dim == 1 else len(hs) * D).astype(np.float32)

        def cin_case(c):
            layer = IL.CIN(conv_size=list(hs), output_dim=out_dim)
            xt = c.t(x0, grad=True)
            layer(xt)
            for conv, w, b in zip(layer.hidden_conv, ws, bs):
                conv.set_weights([w, b])
            if out_dim == 1:
                layer.logit_layer.set_weights([lw, lb])
            y = layer(xt)
            (y * c.t(g)).sum().backward()
            res = {"in/x0": x0, "in/gy": g, "out/y": y, "grad/x0": xt.grad,
                   "meta/seed": np.array(seed), "meta/hs": np.array(hs)}
            for i, conv in enumerate(layer.hidden_conv):
                if store_w:
                    res["w/cin_w%d" % i], res["w/cin_b%d" % i] = ws[i], bs[i]
                    res["grad/cin_w%d" % i] = conv.kernel.grad
                else:                                           # big case: weights are re-drawn from the seed
                    res["grad/cin_w%d_head" % i] = conv.kernel.grad[:, :64]
                    res["grad/cin_w%d_colsum" % i] = conv.kernel.grad.sum(dim=1)
                res["grad/cin_b%d" % i] = conv.bias.grad
            if out_dim == 1:
                res["w/cin_logit_w"], res["w/cin_logit_b"] = lw, lb
                res["grad/cin_logit_w"] = layer.logit_layer.kernel.grad
            return res
        return cin_case
    cases["cin"] = both(ns, make_cin_case(19, 5, 4, (6, 7, 3), 77, True))
    cases["cin_pooled"] = both(ns, make_cin_case(19, 5, 4, (6, 7), 78, True, out_dim=0))
    # the tcgen05 shape: 26 fields, D=16, 200 feature maps (weights re-drawn from meta/seed by the tests)
    cases["cin26"] = both(ns, make_cin_case(24, 26, 16, (200, 200, 200), 79, False))

    # ---- a9: ProductAttentionLayer (BL:292-311) incl. both mask modes ----------------------------
    H, Fa, d = 2, 9, 8
    q = rs.randn(H, B, Fa, d).astype(np.float32)
    kk = rs.randn(H, B, Fa, d).astype(np.float32)
    vv = rs.randn(H, B, Fa, d).astype(np.float32)
    m1 = (rs.rand(Fa, Fa) > 0.3).astype(np.float32)            # mask_mod 1: score @ float(mask)
    m2 = np.triu(np.ones((Fa, Fa)), 0) == 0                    # mask_mod 2 (SeqFM DynamicViewMask, MD:282-289)
    ga = rs.randn(H, B, Fa, d).astype(np.float32)

    def pattn_case(c):
        res = {"in/q": q, "in/k": kk, "in/v": vv, "in/mask1": m1, "in/mask2": m2.astype(np.uint8), "in/gy": ga}
        for tag, kw, mask in (("plain", dict(), None), ("scale", dict(use_scale=True), None),
                              ("mask1", dict(use_scale=True, mask_mod=1), c.t(m1)),
                              ("mask2", dict(use_scale=True, mask_mod=2), torch.as_tensor(m2))):
            qt, kt, vt = c.t(q, grad=True), c.t(kk, grad=True), c.t(vv, grad=True)
            y = BL.ProductAttentionLayer(**kw)([qt, kt, vt], mask=mask)
            (y * c.t(ga)).sum().backward()
            res["out/" + tag] = y
            res["grad/%s_q" % tag], res["grad/%s_k" % tag], res["grad/%s_v" % tag] = qt.grad, kt.grad, vt.grad
        return res
    cases["product_attention"] = both(ns, pattn_case)

    # ---- a10: MultHeadAttentionLayer (BL:335-377) and the DnnLayer wrap (CL:201-226) -------------
    def mha_case_factory(Bm, Fm, k_in, Hh, dd, seed, mask_kind=None):
        r = np.random.RandomState(seed)
        x = r.randn(Bm, Fm, k_in).astype(np.float32)
        wq, wk, wv, wr = [(r.randn(k_in, Hh, dd) * 0.4).astype(np.float32) for _ in range(4)]
        gam = (1 + 0.2 * r.randn(dd)).astype(np.float32)
        bet = (0.1 * r.randn(dd)).astype(np.float32)
        g = r.randn(Hh, Bm, Fm, dd).astype(np.float32)
        mask = None
        if mask_kind == 2:
            mask = np.triu(np.ones((Fm, Fm)), 0) == 0

        def mha_case(c):
            mod = 2 if mask_kind == 2 else 1
            layer = BL.MultHeadAttentionLayer(attention_dim=dd, attention_head_dim=Hh, use_ln=True, atten_mask_mod=mod)
            xt = c.t(x, grad=True)
            mk = None if mask is None else torch.as_tensor(mask)
            layer(xt, mask=mk)
            set_weights([layer.query_w, layer.key_w, layer.value_w, layer.res_w], [wq, wk, wv, wr])
            layer.ln.set_weights([gam, bet])
            out = layer(xt, mask=mk)
            res = {"in/x": x, "in/gy": g, "w/query_w": wq, "w/key_w": wk, "w/value_w": wv, "w/res_w": wr,
                   "w/ln_gamma": gam, "w/ln_beta": bet}
            if mask is not None:
                res["in/mask"] = mask.astype(np.uint8)
            if Hh == 1:
                res["out/atten_v"] = out                       # squeezed [B,F,d] (BL:374-375)
                (out * c.t(g)[0]).sum().backward()
            else:
                res["out/atten_v"], res["out/res"] = out[0], out[1]
                # the block DnnLayer(res_unit=1, other_dense=[layer]) makes of it (AutoInt, MD:160-161)
                xt2 = c.t(x, grad=True)
                blk = CL.DnnLayer(res_unit=1, other_dense=[layer])(xt2)
                res["out/block"] = blk
                for w in (layer.query_w, layer.key_w, layer.res_w, layer.ln.gamma, layer.ln.beta, layer.value_w):
                    w.grad = None
                (blk * c.t(g)).sum().backward()
                res["grad/x"] = xt2.grad
                res["grad/query_w"], res["grad/key_w"], res["grad/res_w"] = layer.query_w.grad, layer.key_w.grad, layer.res_w.grad
                res["grad/ln_gamma"], res["grad/ln_beta"] = layer.ln.gamma.grad, layer.ln.beta.grad
                assert layer.value_w.grad is None              # value_w is never read (BL:360)
            return res
        return mha_case
    cases["mha"] = both(ns, mha_case_factory(21, 26, 16, 2, 8, 501))          # config 5: 2 heads, d=8
    cases["mha_h3"] = both(ns, mha_case_factory(21, 26, 8, 3, 8, 502))        # AutoInt defaults (MD:150)
    cases["mha_h1"] = both(ns, mha_case_factory(21, 7, 8, 1, 8, 503))         # 1 head -> squeezed output
    cases["mha_mask2"] = both(ns, mha_case_factory(21, 7, 8, 2, 8, 504, mask_kind=2))   # SeqFM use (MD:292,296)

    # ---- a12: DnnLayer (CL:159-226) incl. the residual that fires when shapes agree, heads ------
    xin = rs.randn(B, 20).astype(np.float32)
    hu = [12, 12, 5]
    dws, dbs, dprev = [], [], 20
    for u in hu:
        dws.append((rs.randn(dprev, u) * 0.3).astype(np.float32))
        dbs.append((rs.randn(u) * 0.1).astype(np.float32))
        dprev = u
    dlw, dlb = (rs.randn(5, 1) * 0.3).astype(np.float32), (rs.randn(1) * 0.1).astype(np.float32)

    def dnn_case(c):
        layer = CL.DnnLayer(hidden_units=list(hu), output_dim=1)
        xt = c.t(xin, grad=True)
        layer(xt)
        for h, w, b in zip(layer.hidden_list, dws, dbs):
            h.dense.set_weights([w, b])
        layer.logit_layer.set_weights([dlw, dlb])
        y = layer(xt)
        y.sum().backward()
        res = {"in/x": xin, "out/y": y, "grad/x": xt.grad, "w/dnn_logit_w": dlw, "w/dnn_logit_b": dlb}
        for i in range(len(hu)):
            res["w/dnn_w%d" % i], res["w/dnn_b%d" % i] = dws[i], dbs[i]
            res["grad/dnn_w%d" % i] = layer.hidden_list[i].dense.kernel.grad
        return res
    cases["dnn"] = both(ns, dnn_case)

    x1 = rs.randn(B, 1, 8).astype(np.float32)
    x2 = rs.randn(B, 6).astype(np.float32)
    hw, hb = (rs.randn(14, 2) * 0.3).astype(np.float32), (rs.randn(2) * 0.1).astype(np.float32)
    s1, s2, s3 = rs.randn(B, 1, 1).astype(np.float32), rs.randn(B, 1).astype(np.float32), rs.randn(B, 1).astype(np.float32)

    def heads_case(c):
        layer = CL.MergeScoreLayer()
        a, b = c.t(x1), c.t(x2)
        layer([a, b])
        layer.dense.set_weights([hw, hb])
        return {"in/x1": x1, "in/x2": x2, "w/head_w": hw, "w/head_b": hb, "out/merge": layer([a, b]),
                "in/s1": s1, "in/s2": s2, "in/s3": s3,
                "out/score_add": CL.ScoreLayer(use_add=True)([c.t(s1), c.t(s2), c.t(s3)]),
                "out/score": CL.ScoreLayer()(c.t(s1))}
    cases["heads"] = both(ns, heads_case)
    # ---- DP:85-102, 294-301: the reference's own feature encoders (pandas + sklearn, no TensorFlow involved) ----
    import pandas as pd
    r2 = np.random.RandomState(77)
    n_rows = 60
    cats = np.array(["a", "b", "10", "9", "zz", "A", "-1", "x y"])
    sdf = pd.DataFrame({"C14": cats[r2.randint(0, 8, n_rows)], "C15": r2.randint(0, 5, n_rows).astype(object),
                        "C16": cats[r2.randint(0, 3, n_rows)]})
    sdf.loc[[3, 17, 40], "C14"] = np.nan
    sdf.loc[[5, 6], "C15"] = np.nan
    ddf = pd.DataFrame({"I1": r2.rand(n_rows) * 50 - 10, "I2": r2.randint(0, 4, n_rows).astype(float), "I3": np.full(n_rows, 2.5)})
    ddf.loc[[1, 2, 30], "I1"] = np.nan
    ddf.loc[[9], "I2"] = np.nan
    dp2 = DP.data_prepare(batch_size=16)
    enc, sinfo = dp2.sparse_fea_deal(sdf.copy())
    den, dinfo = dp2.dense_fea_deal(ddf.copy())
    cases["data_prepare"] = {
        "in/sparse": np.array(sdf.fillna("<nan>").astype(str).to_numpy(), dtype="U8"),
        "in/dense": ddf.to_numpy(dtype=np.float64),
        "out/ids": enc.to_numpy().astype(np.int64), "out/word_size": np.array([i.word_size for i in sinfo]),
        "out/cross_unit": np.array([i.cross_unit for i in sinfo]), "out/emb_reg": np.array([i.emb_reg for i in sinfo]),
        "out/dense": den.to_numpy(dtype=np.float64),
        "out/fields": np.array(list(sinfo[0]._fields), dtype="U16"),
    }
    return cases


# ------------------------------------------------------------------------------------------------
# model cases: the reference's builders, run end to end (FeatureInput -> builder -> loss -> grads)
# ------------------------------------------------------------------------------------------------
def _extract_params(ns, model):
    """Reference layer objects -> {oracle name: weight tensor} (names of oracle/kon_oracle.py)."""
    IL, CL, BL = ns.IL, ns.CL, ns.BL
    p = {}
    n_logit = 0
    for layer in model.layers:
        if isinstance(layer, IL.SparseEmbed):
            pre = "lin_" if layer.is_linear else "emb_"
            for f, e in enumerate(layer.embed):
                p[pre + str(f)] = e.embeddings
        elif isinstance(layer, CL.DnnLayer):
            if layer.hidden_list and isinstance(layer.hidden_list[0], CL.HiddenLayer):
                for i, h in enumerate(layer.hidden_list):
                    p["dnn_w%d" % i], p["dnn_b%d" % i] = h.dense.kernel, h.dense.bias
            if layer.output_dim != -1:
                p["dnn_logit_w"], p["dnn_logit_b"] = layer.logit_layer.kernel, layer.logit_layer.bias
        elif isinstance(layer, CL.MergeScoreLayer):
            p["head_w"], p["head_b"] = layer.dense.kernel, layer.dense.bias
        elif isinstance(layer, IL.CrossLayer):
            for i in range(layer.cross_hidden):
                p["outer_weight_%d" % i], p["outer_bias_%d" % i] = layer.kernel[i], layer.bias[i]
        elif isinstance(layer, IL.CIN):
            for i, conv in enumerate(layer.hidden_conv):
                p["cin_w%d" % i], p["cin_b%d" % i] = conv.kernel, conv.bias
            if layer.output_dim == 1:
                p["cin_logit_w"], p["cin_logit_b"] = layer.logit_layer.kernel, layer.logit_layer.bias
        elif isinstance(layer, BL.MultHeadAttentionLayer):
            p["query_w"], p["key_w"], p["value_w"], p["res_w"] = layer.query_w, layer.key_w, layer.value_w, layer.res_w
            p["ln_gamma"], p["ln_beta"] = layer.ln.gamma, layer.ln.beta
        elif isinstance(layer, IL.AttentionBaseLayer):
            p["afm_score_w"], p["afm_score_b"] = layer.kernel_w, layer.kernel_b
            p["afm_mlp_w"] = layer.single_mlp.kernel
            p["afm_out_w"], p["afm_out_b"] = layer.output_layer.kernel, layer.output_layer.bias
    del n_logit
    return p


def model_cases(ns):
    DP, MD, tf = ns.DP, ns.MD, ns.tf
    cases = {}
    F = 26

    def run_model(builder, k, B, seed, useAddLinear=False, builder_kw=None, sigmoid=False, scale_hook=True):
        rs = np.random.RandomState(seed)
        rows = [3 + (7 * f) % 23 for f in range(F)]
        ids = np.stack([rs.randint(0, r, B) for r in rows], 1).astype(np.float32)
        dense = rs.rand(B, N_DENSE).astype(np.float32)
        y = (rs.rand(B) < 0.3).astype(np.int64)
        labels = (y.astype(np.float32).reshape(B, 1, 1) if sigmoid else
                  np.stack([1 - y, y], 1).astype(np.float32))                   # to_categorical, DP:359

        def fn(c):
            dp = DP.data_prepare(batch_size=None)
            sinfo = [dp.sparseFea(fea_name="C%d" % (f + 14), word_size=rows[f], input_dim=B, cross_unit=k,
                                  linear_unit=1, pre_weight=None, mask_zero=False, is_trainable=True,
                                  input_length=1, sample_num=None, batch_size=None, emb_reg=1e-8) for f in range(F)]
            dinfo = [dp.denseFea("I%d" % (j + 1), None) for j in range(N_DENSE)]
            feed = {"C%d" % (f + 14): ids[:, f:f + 1] for f in range(F)}
            feed.update({"I%d" % (j + 1): dense[:, j:j + 1] for j in range(N_DENSE)})
            tf.keras.feed(feed)
            if scale_hook:
                # Keras' default inits give near-zero logits; widen the tables / biases so that every
                # term of the model matters in the fixture (weights are stored, so this is only data)
                def hook(w):
                    if w.kon_name.endswith("embeddings"):
                        return w * (0.6 if w.shape[1] > 1 else 6.0)
                    if w.kon_name.endswith("bias") and w.abs().max() == 0:
                        g = torch.Generator().manual_seed(int(w.numel()) + 17)
                        return (torch.randn(w.shape, generator=g, dtype=torch.float64) * 0.05).to(w.dtype)
                    if w.kon_name.endswith("outer_bias_0") or "outer_bias" in w.kon_name:
                        g = torch.Generator().manual_seed(int(w.numel()) + 19)
                        return (torch.randn(w.shape, generator=g, dtype=torch.float64) * 0.05).to(w.dtype)
                    if w.kon_name.endswith("gamma"):
                        g = torch.Generator().manual_seed(int(w.numel()) + 23)
                        return (1 + 0.2 * torch.randn(w.shape, generator=g, dtype=torch.float64)).to(w.dtype)
                    return None
                tf.keras.initializers.set_hook(hook)
            inp = dp.FeatureInput(sparseInfo=sinfo, denseInfo=dinfo, useLinear=True, useAddLinear=useAddLinear)
            model = builder(inp, **(builder_kw or {}))
            out = model.outputs
            p = _extract_params(ns, model)
            loss = tf.keras.losses.binary_crossentropy(c.t(labels), out).mean()
            names = [n for n in p if p[n].requires_grad]
            grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
            res = {"in/ids": ids.astype(np.int32), "in/dense": dense, "in/labels": labels,
                   "meta/rows": np.array(rows), "meta/k": np.array(k), "out/y": out, "out/loss": loss}
            for n in p:
                res["w/" + n] = p[n]
            for n, g in zip(names, grads):
                if g is not None:
                    res["grad/" + n] = g
            # Keras adds the regulariser penalties of the model to the training loss (IL:217)
            reg = sum((w.kon_regularizer(w) for w in model.weights if getattr(w, "kon_regularizer", None) is not None),
                      torch.zeros((), dtype=c.dt))
            res["out/reg_loss"] = reg
            return res
        return both(ns, fn)

    cases["fm"] = run_model(MD.FM, 8, 48, 1)
    cases["deepfm"] = run_model(MD.DeepFM, 16, 48, 2, builder_kw=dict(hidden_units=[48, 24, 12]))
    cases["dcn"] = run_model(MD.DCN, 8, 48, 3, builder_kw=dict(cross_hidden=6, hidden_units=[32, 16, 16]))
    cases["xdeepfm"] = run_model(MD.XDeepFM, 16, 40, 4, useAddLinear=True, sigmoid=True,
                                 builder_kw=dict(conv_size=[10, 9, 8], hidden_units=[40, 20, 10]))
    cases["autoint"] = run_model(MD.AutoInt, 16, 48, 5, builder_kw=dict(attention_dim=8, attention_head_dim=2))
    cases["nfm"] = run_model(MD.NFM, 8, 48, 6, sigmoid=True, builder_kw=dict(hidden_units=[32, 16, 8]))
    cases["afm"] = run_model(MD.AFM, 8, 48, 7, sigmoid=True)
    # IPNN (MD:43-56 with use_outer=False: OPnnLayer is broken in the reference, IL:56 vs IL:63)
    cases["pnn"] = run_model(MD.PNN, 8, 48, 8, builder_kw=dict(hidden_units=[32, 32, 16], use_inner=True, use_outer=False))
    return cases


def generate():
    ns = run_reference.load()
    flat_l, flat_m = {}, {}
    for cname, d in layer_cases(ns).items():
        for k, v in d.items():
            flat_l["%s/%s" % (cname, k)] = v
    for cname, d in model_cases(ns).items():
        for k, v in d.items():
            flat_m["%s/%s" % (cname, k)] = v
    return flat_l, flat_m


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def check(verbose=True):
    """Regenerate from /root/reference and compare with the committed files, bit for bit."""
    fl, fm = generate()
    bad = []
    for path, fresh in ((LAYERS_NPZ, fl), (MODELS_NPZ, fm)):
        stored = np.load(path)
        if set(stored.files) != set(fresh):
            bad.append("%s: key sets differ (%d stored, %d fresh)" % (os.path.basename(path), len(stored.files), len(fresh)))
            continue
        for k in stored.files:
            if not _same(stored[k], fresh[k]):
                bad.append("%s:%s differs" % (os.path.basename(path), k))
    if verbose:
        print("ref golden check:", "OK" if not bad else bad[:10])
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        sys.exit(1 if check() else 0)
    fl, fm = generate()
    np.savez_compressed(LAYERS_NPZ, **fl)
    np.savez_compressed(MODELS_NPZ, **fm)
    for p, d in ((LAYERS_NPZ, fl), (MODELS_NPZ, fm)):
        print("%s: %d arrays, %.1f kB" % (os.path.relpath(p, ROOT), len(d), os.path.getsize(p) / 1024))


if __name__ == "__main__":
    main()
