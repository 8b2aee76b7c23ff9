"""-m gpu: the committed golden vectors and whole models (reference builders) through the
C-ABI, forward and backward, against the CPU oracle with identical injected weights."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_rel, gen, offsets, rel_err
from oracle import kon_oracle as ko

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kon_golden.npz"))


def _g(name):
    return torch.from_numpy(GOLD[name]).to(DEV)


def test_golden_vectors_through_the_abi():
    from ml_function_b200 import _lib as L, ops
    offs = offsets(GOLD["emb_rows"].tolist())
    ids = _g("emb_ids")
    out = ops.embed_fwd_raw(_g("emb_tables"), ids, offs)
    assert torch.equal(out.cpu(), torch.from_numpy(GOLD["emb_out"]))               # bit-exact lookup
    sg = ops.embed_bwd_raw(_g("emb_dout"), ids, offs)
    n = int(sg.n.item())
    assert np.array_equal(sg.rows[:n].cpu().numpy(), GOLD["emb_unique_rows"])       # bit-exact routing
    assert_rel(sg.grads[:n], torch.from_numpy(GOLD["emb_grads"]), 1e-6, "golden emb grads")
    lin = ops.embed_fwd_raw(_g("emb_lins"), ids, offs)
    assert_rel(ops.fm(out, lin[..., 0]), torch.from_numpy(GOLD["fm_out_f64"]), 1e-5, "golden fm")
    cw, cb = _g("cross_w")[..., 0].contiguous(), _g("cross_b")[..., 0].contiguous()
    assert_rel(ops.cross(_g("cross_x"), cw, cb), torch.from_numpy(GOLD["cross_out"][..., 0]), 1e-5, "golden cross")
    pooled = ops.cin(_g("cin_x0"), [_g("cin_w0")[0], _g("cin_w1")[0]], [_g("cin_b0"), _g("cin_b1")], L.KON_CIN_FP32)
    assert_rel(pooled, torch.from_numpy(GOLD["cin_pooled_f64"]), 1e-5, "golden cin")
    y = ops.attention(_g("attn_x"), _g("attn_wq"), _g("attn_wk"), _g("attn_wr"), _g("attn_gamma"), _g("attn_beta"))
    assert_rel(y, torch.from_numpy(GOLD["attn_out"]), 1e-5, "golden attention")


# ---------------------------------------------------------------------------------------
def _params(name, rows, k, g, hidden=(24, 16, 8), conv=(10, 9, 8), heads=2, d=4, cross=3):
    F = len(rows)
    p = {}
    for f, r in enumerate(rows):
        p[f"emb_{f}"] = torch.randn(r, k, generator=g) * 0.5
        p[f"lin_{f}"] = torch.randn(r, 1, generator=g) * 0.5
    D = 13 + F * k
    dims = [D] + list(hidden)
    for i in range(3):
        p[f"dnn_w{i}"] = ko.glorot_uniform((dims[i], dims[i + 1]), g)
        p[f"dnn_b{i}"] = torch.randn(dims[i + 1], generator=g) * 0.1
    if name == "fm":
        p["head_w"], p["head_b"] = ko.glorot_uniform((k, 2), g), torch.randn(2, generator=g) * 0.1
    if name == "deepfm":
        p["head_w"], p["head_b"] = ko.glorot_uniform((k + hidden[-1], 2), g), torch.randn(2, generator=g) * 0.1
    if name == "dcn":
        for i in range(cross):
            p[f"outer_weight_{i}"] = ko.glorot_uniform((D, 1), g)
            p[f"outer_bias_{i}"] = torch.randn(D, 1, generator=g) * 0.05
        p["head_w"], p["head_b"] = ko.glorot_uniform((D + hidden[-1], 2), g), torch.zeros(2)
    if name == "xdeepfm":
        hp = F
        for i, n in enumerate(conv):
            p[f"cin_w{i}"] = ko.glorot_uniform((1, hp * F, n), g)
            p[f"cin_b{i}"] = torch.randn(n, generator=g) * 0.05
            hp = n
        p["cin_logit_w"], p["cin_logit_b"] = ko.glorot_uniform((len(conv) * k, 1), g), torch.zeros(1)
        p["dnn_logit_w"], p["dnn_logit_b"] = ko.glorot_uniform((hidden[-1], 1), g), torch.zeros(1)
    if name == "nfm":
        dims = [13 + k] + list(hidden)
        for i in range(3):
            p[f"dnn_w{i}"] = ko.glorot_uniform((dims[i], dims[i + 1]), g)
        p["dnn_logit_w"], p["dnn_logit_b"] = ko.glorot_uniform((hidden[-1], 1), g), torch.zeros(1)
    if name == "autoint":
        for w in ("query_w", "key_w", "res_w"):
            p[w] = ko.glorot_uniform((k, heads, d), g)
        p["ln_gamma"], p["ln_beta"] = torch.rand(d, generator=g) + 0.5, torch.randn(d, generator=g) * 0.1
        p["head_w"], p["head_b"] = ko.glorot_uniform((heads * F * d, 2), g), torch.zeros(2)
    return p


def _build(name, rows, k, p, cin_precision="fp32", hidden=(24, 16, 8), conv=(10, 9, 8)):
    from ml_function_b200 import layers as KL, models as KM
    sp = [KL.make_sparse_fea(str(14 + i), r, cross_unit=k) for i, r in enumerate(rows)]
    de = [KL.denseFea(str(1 + i), None) for i in range(13)]
    fea = KM.FeatureInput(sp, de, useLinear=True, useAddLinear=(name == "xdeepfm"), device=DEV)
    m = {"fm": lambda: KM.FM(fea), "deepfm": lambda: KM.DeepFM(fea, hidden_units=list(hidden)),
         "dcn": lambda: KM.DCN(fea, hidden_units=list(hidden), cross_hidden=3),
         "xdeepfm": lambda: KM.XDeepFM(fea, conv_size=list(conv), hidden_units=list(hidden), cin_precision=cin_precision),
         "nfm": lambda: KM.NFM(fea, hidden_units=list(hidden)),
         "autoint": lambda: KM.AutoInt(fea, attention_dim=4, attention_head_dim=2)}[name]()
    m.load_reference_params({k_: v.to(DEV) for k_, v in p.items()})
    return m


ORACLE = {"fm": ko.model_fm, "deepfm": ko.model_deepfm, "dcn": ko.model_dcn, "xdeepfm": ko.model_xdeepfm,
          "autoint": ko.model_autoint, "nfm": ko.model_nfm}


@pytest.mark.parametrize("name", ["fm", "deepfm", "dcn", "xdeepfm", "autoint", "nfm"])
def test_model_forward_backward_matches_oracle(name):
    from ml_function_b200.models import keras_binary_crossentropy
    g = gen(11)
    rows = [7, 300, 5, 41, 2, 1000]
    B, k = 517, 8
    p = _params(name, rows, k, g)
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
    dense = torch.rand(B, 13, generator=g)
    y = (torch.rand(B, generator=g) < 0.3).float()
    labels = y.view(B, 1, 1) if name in ("xdeepfm", "nfm") else torch.stack([1 - y, y], 1)
    # oracle in fp64 = truth; in fp32 = the reference's own arithmetic (its error sets the scale)
    p64 = {k_: v.double().requires_grad_(True) for k_, v in p.items()}
    out64 = ORACLE[name](p64, dense.double(), ids)
    loss64 = ko.binary_crossentropy(labels.double(), out64)
    loss64.backward()
    out32 = ORACLE[name](p, dense, ids)
    model = _build(name, rows, k, p)
    out = model(dense.to(DEV), ids.to(DEV))
    assert out.shape == out64.shape
    e_ref = rel_err(out32, out64)
    e = assert_rel(out, out64, max(1e-5, 4 * e_ref), f"{name} forward")
    loss = keras_binary_crossentropy(labels.to(DEV), out)
    loss.backward()
    assert abs(loss.item() - loss64.item()) < 1e-5 * max(1.0, abs(loss64.item()))
    # sparse gradient of the embedding arena == dense oracle gradient, row for row
    offs = offsets(rows)
    sgs = model.sparse_embed.arena.kon_sparse_grads
    assert len(sgs) == 1 and model.sparse_embed.arena.grad is None
    dense_g = sgs[0].to_dense(offs[-1]).cpu()
    ref_g = torch.cat([p64[f"emb_{f}"].grad for f in range(len(rows))])
    assert_rel(dense_g, ref_g, 2e-5, f"{name} embedding grad")
    n = int(sgs[0].n.item())
    touched = torch.unique((ids.long() + torch.tensor(offs[:-1])).reshape(-1))
    assert torch.equal(sgs[0].rows[:n].cpu().long(), touched)                       # bit-exact routing
    if name in ("fm", "deepfm", "xdeepfm", "nfm"):
        lg = model.linear_embed.arena.kon_sparse_grads[0].to_dense(offs[-1]).cpu()
        assert_rel(lg, torch.cat([p64[f"lin_{f}"].grad for f in range(len(rows))]), 2e-5, f"{name} linear grad")
    # a few dense weights
    if name == "dcn":
        ref_w = torch.stack([model.ref_to_phys_rows(p64[f"outer_weight_{i}"].grad)[:, 0] for i in range(3)])
        assert_rel(model.cross.kernel.grad, ref_w, 2e-5, "dcn cross kernel grad")
    if name == "xdeepfm":
        for i in range(3):
            assert_rel(model.cin.conv_kernels[i].grad, p64[f"cin_w{i}"].grad, 2e-5, f"cin_w{i} grad")
            assert_rel(model.cin.conv_biases[i].grad, p64[f"cin_b{i}"].grad, 2e-5, f"cin_b{i} grad")
    if name == "autoint":
        att = model.blocks[0].other_dense[0]
        assert_rel(att.query_w.grad, p64["query_w"].grad, 2e-5, "query_w grad")
        assert_rel(att.key_w.grad, p64["key_w"].grad, 2e-5, "key_w grad")
        assert att.value_w.grad is None                                             # never read (BL:360)
    if name in ("deepfm", "dcn", "xdeepfm"):
        assert_rel(model.dnn.kernels[0].grad, model.ref_to_phys_rows(p64["dnn_w0"].grad), 2e-5, "dnn_w0 grad")


def test_reference_list_call_convention():
    """Model called the way the reference calls it: lists of per-feature [B,1] tensors, ids as
    float32 (DP:290-292)."""
    g = gen(5)
    rows = [9, 30, 4]
    B, k = 64, 8
    p = _params("deepfm", rows, k, g)
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1)
    dense = torch.rand(B, 13, generator=g)
    model = _build("deepfm", rows, k, p)
    a = model(dense.to(DEV), ids.to(torch.int32).to(DEV))
    b = model([dense[:, j:j + 1].to(DEV) for j in range(13)], [ids[:, f:f + 1].float().to(DEV) for f in range(3)])
    assert torch.equal(a, b)


def test_train_step_updates_only_touched_rows_and_lowers_loss():
    from ml_function_b200.train import Trainer
    g = gen(9)
    rows = [50, 3000, 7]
    B, k = 1024, 8
    p = _params("deepfm", rows, k, g)
    model = _build("deepfm", rows, k, p)
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).to(DEV)
    dense = torch.rand(B, 13, generator=g).to(DEV)
    y = (torch.rand(B, generator=g) < 0.3).float()
    labels = torch.stack([1 - y, y], 1).to(DEV)
    tr = Trainer(model, lr=1e-2)
    before = model.sparse_embed.arena.detach().clone()
    l0 = tr.step(dense, ids, labels).item()
    for _ in range(20):
        l1 = tr.step(dense, ids, labels).item()
    assert l1 < l0
    changed = (model.sparse_embed.arena.detach() != before).any(dim=1).nonzero().flatten().cpu()
    touched = torch.unique((ids.cpu().long() + torch.tensor(offsets(rows)[:-1])).reshape(-1))
    assert torch.equal(changed, touched)


def test_data_pipeline_feeds_the_trainer_from_hbm_and_pinned_memory():
    """DP:335-337 pipeline (shuffle(2048).repeat(2).batch().prefetch(2)) over a dataset resident in HBM / in pinned
    host memory -> Trainer.step: every sample is seen exactly twice, batches arrive on the device as int32 ids."""
    import numpy as np
    from ml_function_b200.data_prepare import data_prepare
    from ml_function_b200.train import Trainer
    g = gen(13)
    rows = [50, 3000, 7]
    n, k = 3000, 8
    p = _params("deepfm", rows, k, g)
    ids = torch.stack([torch.randint(0, r, (n,), generator=g) for r in rows], 1).numpy().astype(np.int32)
    ids[:, 1] = np.arange(n) % rows[1]                    # field 1 identifies the sample (mod 3000 = identity)
    dense = torch.rand(n, 13, generator=g).numpy()
    y = (torch.rand(n, generator=g) < 0.3).float()
    labels = torch.stack([1 - y, y], 1).numpy()
    for resident in ("device", "host"):
        model = _build("deepfm", rows, k, p)
        tr = Trainer(model, lr=1e-3)
        dp = data_prepare(batch_size=512, device=DEV)
        pipe = dp.data_pipeline(((ids, dense), labels), resident=resident)
        seen, losses = [], []
        for d_, i_, y_ in pipe:
            assert i_.is_cuda and i_.dtype == torch.int32 and d_.is_cuda and y_.shape[1] == 2
            seen.append(i_[:, 1].clone())
            losses.append(tr.step(d_, i_, y_))
        torch.cuda.synchronize()
        counts = torch.bincount(torch.cat(seen).long().cpu(), minlength=n)
        assert bool((counts == 2).all()) and len(losses) == len(pipe) == (2 * n + 511) // 512
        assert all(torch.isfinite(l) for l in losses)


def test_sparse_adam_sums_the_gradients_of_an_arena_looked_up_twice():
    """Keras' optimizer sees ONE gradient per variable (the sum over its uses): a row hit by two lookups of
    the same arena in one step gets one Adam update on the summed gradient, not two updates."""
    from ml_function_b200 import ops
    from ml_function_b200.train import SparseAdam
    g = gen(31)
    rows, k, B = [40, 9], 4, 64
    offs = offsets(rows)
    arena = torch.nn.Parameter(torch.randn(sum(rows), k, generator=g).to(DEV))
    w0 = arena.detach().clone()
    ids_a = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).to(DEV)
    ids_b = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).to(DEV)
    da = torch.randn(B, 2, k, generator=g).to(DEV)
    db = torch.randn(B, 2, k, generator=g).to(DEV)
    arena.kon_sparse_grads = [ops.embed_bwd_raw(da, ids_a, offs), ops.embed_bwd_raw(db, ids_b, offs)]
    opt = SparseAdam(arena, lr=1e-2)
    opt.step()
    # oracle: dense Adam on the summed gradient, rows without a gradient untouched (lazy)
    dense = torch.zeros(sum(rows), k, dtype=torch.float64)
    o = torch.tensor(offs[:-1])
    for ids, d in ((ids_a, da), (ids_b, db)):
        dense.index_add_(0, (ids.cpu().long() + o).reshape(-1), d.cpu().double().reshape(-1, k))
    m = 0.1 * dense
    v = 0.001 * dense * dense
    upd = 1e-2 * (m / (1 - 0.9)) / ((v / (1 - 0.999)).sqrt() + 1e-7)
    want = w0.cpu().double() - torch.where(dense != 0, upd, torch.zeros_like(upd))
    assert_rel(arena.detach().cpu().double(), want, 1e-5)


def test_dcn_alignment_columns_stay_inert_through_training():
    """The concat buffer is padded to a multiple of 4 floats; the cross layer's pad bias must never train
    (else the pad columns become learnt constant features the reference does not have)."""
    from ml_function_b200 import layers as KL, models as KM
    from ml_function_b200.train import Trainer
    g = gen(12)
    rows = [50, 30, 7]
    B, k = 256, 4                                     # 3*4 + 13 = 25 -> W = 28: three alignment columns
    sparse = [KL.make_sparse_fea(f"s{i}", r, cross_unit=k) for i, r in enumerate(rows)]
    dense_info = [KL.denseFea(f"d{i}", None) for i in range(13)]
    fea = KM.FeatureInput(sparse, dense_info, useLinear=True, useAddLinear=True, device=DEV)
    model = KM.DCN(fea, hidden_units=[16, 8], cross_hidden=3).to(DEV)
    nv = model.Fk + model.n_dense
    assert model.W > nv
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).to(DEV)
    dense = torch.rand(B, 13, generator=g).to(DEV)
    y = (torch.rand(B, generator=g) < 0.3).float()
    labels = torch.stack([1 - y, y], 1).to(DEV)
    tr = Trainer(model, lr=1e-2)
    for _ in range(5):
        tr.step(dense, ids, labels)
    assert model.cross.kernel.shape[1] == model.W
    assert torch.count_nonzero(model.cross.bias.detach()[:, nv:]) == 0
    assert torch.count_nonzero(model.cross.kernel.detach()[:, nv:]) == 0
    assert torch.count_nonzero(model.cross.bias.detach()[:, :nv]) > 0


def test_keras_optimizer_mode_matches_the_reference_update_over_several_steps():
    """``Trainer(optimizer="keras")``: the reference's own ``compile(optimizer='adam')`` -- dense Adam on EVERY row of
    the embedding tables incl. the regulariser's dense ``2 l2 w`` term (IL:217) -- against an fp64 restatement of the
    Keras update rule over three steps on different batches.  The default (row-wise lazy) mode provably differs:
    rows no batch touches stay put there, and move here."""
    from ml_function_b200.train import Trainer
    g = gen(77)
    rows = [7, 300, 5, 41, 2, 1000]
    B, k, steps, lr, l2 = 96, 8, 3, 1e-3, 1e-8
    p = _params("deepfm", rows, k, g)
    batches = []
    for _ in range(steps):
        ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
        y = (torch.rand(B, generator=g) < 0.3).float()
        batches.append((torch.rand(B, 13, generator=g), ids, torch.stack([1 - y, y], 1)))
    # fp64 truth: autograd over the oracle's DeepFM + the Keras Adam formulas on every weight
    w = {n: v.double().clone().requires_grad_(True) for n, v in p.items()}
    m = {n: torch.zeros_like(v) for n, v in w.items()}
    vv = {n: torch.zeros_like(v) for n, v in w.items()}
    b1, b2, eps = 0.9, 0.999, 1e-7
    for t, (dense, ids, labels) in enumerate(batches, 1):
        loss = ko.binary_crossentropy(labels.double(), ko.model_deepfm(w, dense.double(), ids))
        grads = torch.autograd.grad(loss, list(w.values()), allow_unused=True)
        lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        with torch.no_grad():
            for (n, wt), gr in zip(w.items(), grads):
                gr = torch.zeros_like(wt) if gr is None else gr.clone()
                if n.startswith("emb_"):
                    gr += 2 * l2 * wt
                m[n].mul_(b1).add_(gr, alpha=1 - b1)
                vv[n].mul_(b2).addcmul_(gr, gr, value=1 - b2)
                wt -= lr_t * m[n] / (vv[n].sqrt() + eps)
    model = _build("deepfm", rows, k, p)
    tr = Trainer(model, lr=lr, optimizer="keras")
    for dense, ids, labels in batches:
        tr.step(dense.to(DEV), ids.to(DEV), labels.to(DEV))
    assert tr.capture(*[t.to(DEV) for t in batches[0]]) is False        # host-side step count: eager only
    off = offsets(rows)
    arena, lin = model.sparse_embed.arena.detach().cpu().double(), model.linear_embed.arena.detach().cpu().double()
    tol = 2e-5
    for f in range(len(rows)):
        assert (arena[off[f]:off[f + 1]] - w[f"emb_{f}"].detach()).abs().max() < tol, f
        assert (lin[off[f]:off[f + 1]] - w[f"lin_{f}"].detach()).abs().max() < tol, f
    k0 = model.phys_to_ref_rows(model.dnn.kernels[0].detach()).cpu().double()
    assert (k0 - w["dnn_w0"].detach()).abs().max() < tol
    for i in (1, 2):
        assert (model.dnn.kernels[i].detach().cpu().double() - w[f"dnn_w{i}"].detach()).abs().max() < tol
    for i in range(3):
        assert (model.dnn.biases[i].detach().cpu().double() - w[f"dnn_b{i}"].detach()).abs().max() < tol
    # a row of the 1000-row table that no batch looked up: moved by the dense regulariser / Adam here ...
    seen = torch.cat([b[1][:, 5] for b in batches]).unique()
    untouched = next(r for r in range(1000) if r not in set(seen.tolist()))
    moved = (arena[off[5] + untouched] - p["emb_5"][untouched].double()).abs().max()
    assert moved > 1e-7
    # ... and left alone by the default lazy mode
    lazy = _build("deepfm", rows, k, p)
    tl = Trainer(lazy, lr=lr)
    for dense, ids, labels in batches:
        tl.step(dense.to(DEV), ids.to(DEV), labels.to(DEV))
    assert torch.equal(lazy.sparse_embed.arena.detach().cpu()[off[5] + untouched], p["emb_5"][untouched])


@pytest.mark.parametrize("name", ["deepfm", "dcn"])
def test_concat_buffer_gradient_accumulated_in_place_equals_autograd_sum(name):
    """Inside a Trainer step the FM / cross backward adds its gradient of the concat buffer INTO the first Dense
    layer's input gradient (kon_fm_bwd_acc / kon_cross_bwd_acc) instead of letting autograd sum two [B,W] tensors:
    same weights after the step as with the fusion switched off, and the fused path is really the one taken."""
    from ml_function_b200 import ops
    from ml_function_b200.train import Trainer
    g = gen(5)
    rows = [7, 300, 5, 41, 2, 1000]
    B, k = 257, 8
    p = _params(name, rows, k, g)
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).to(DEV)
    dense = torch.rand(B, 13, generator=g).to(DEV)
    y = (torch.rand(B, generator=g) < 0.3).float()
    labels = torch.stack([1 - y, y], 1).to(DEV)
    res, hits = {}, {}
    real_take = ops.take_xgrad
    for mode in (True, False):
        ops.ACC_XGRAD = mode
        n = [0]

        def counting(key, shape, _n=n):
            r = real_take(key, shape)
            _n[0] += r is not None
            return r
        ops.take_xgrad = counting
        try:
            model = _build(name, rows, k, p)
            tr = Trainer(model, lr=1e-2)
            tr.step(dense, ids, labels)
            tr.step(dense, ids, labels)
        finally:
            ops.take_xgrad = real_take
            ops.ACC_XGRAD = True
        hits[mode] = n[0]
        res[mode] = [model.sparse_embed.arena.detach().clone(), model.linear_embed.arena.detach().clone()] + \
                    [w.detach().clone() for w in model.dnn.kernels]
    assert hits[True] == 2 and hits[False] == 0          # one fused accumulation per step
    for a, b in zip(res[True], res[False]):
        assert_rel(a, b.double(), 1e-6, name + " weights after two steps, fused vs autograd sum")
