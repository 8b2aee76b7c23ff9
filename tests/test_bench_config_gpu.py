"""-m gpu: parity of EXACTLY what bench.py measures.

* the CUDA-graph replay of the training step (Trainer.capture / step_graph) leaves the same weights and
  optimizer state as eager steps;
* xDeepFM at the benched configuration -- B = 65,536, full Criteo tables, bf16 tcgen05 CIN [200,200,200]
  (incl. the last-layer pooled shortcut), bf16 MLP -- against the fp64 oracle on sampled rows: per-sample
  outputs and per-sample input gradients (rows are independent given the weights), tolerance 2e-2;
* the 3-layer x 2-head bf16 AutoInt stack (the benched AutoInt) against the stacked oracle.
"""
import pytest
import torch

import bench
from helpers import assert_rel, gen, rel_err
from oracle import kon_oracle as ko

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SMALL_ROWS = [3 + (11 * f) % 97 for f in range(26)]


def _batches(n, B, rows, sigmoid, seed=5):
    return [tuple(t.to(DEV) for t in b) for b in bench.synth_batches(n, B, rows, seed, sigmoid, pin=False)]


@pytest.mark.parametrize("name", ["xdeepfm", "deepfm", "autoint"])
def test_step_graph_leaves_the_same_state_as_eager_steps(name):
    from ml_function_b200.train import Trainer
    B, K = 512, 4
    bs = _batches(3, B, SMALL_ROWS, name == "xdeepfm")

    def make():
        torch.manual_seed(2020)
        m = bench.build_model(name, DEV, cin_precision="bf16", rows=SMALL_ROWS, mlp_dtype=torch.bfloat16)
        return m, Trainer(m, lr=1e-2)
    m_e, t_e = make()
    m_g, t_g = make()
    for _ in range(3):                       # what capture() runs as warm-up, on the same batch
        t_e.step(*bs[0])
    assert t_g.capture(*bs[0], warmup=3), t_g.capture_error
    for i in range(K):
        le = t_e.step(*bs[i % 3])
        lg = t_g.step_graph(*bs[i % 3]).clone()
    torch.cuda.synchronize()
    assert torch.equal(le, lg), (float(le), float(lg))
    pe, pg = dict(m_e.named_parameters()), dict(m_g.named_parameters())
    assert pe.keys() == pg.keys() and len(pe) >= 4
    for n_ in pe:
        assert torch.equal(pe[n_].detach(), pg[n_].detach()), f"{name}: parameter {n_} differs after {K} steps"
    for so_e, so_g in zip(t_e.sparse_opts, t_g.sparse_opts):
        assert torch.equal(so_e.m, so_g.m) and torch.equal(so_e.v, so_g.v) and int(so_e.t) == int(so_g.t)
    for (k_e, st_e), (k_g, st_g) in zip(t_e.dense_opt.state.items(), t_g.dense_opt.state.items()):
        assert torch.equal(st_e["exp_avg"], st_g["exp_avg"]) and torch.equal(st_e["exp_avg_sq"], st_g["exp_avg_sq"])
    t_g.release_graph()


def test_xdeepfm_at_the_benched_config_sampled_rows():
    from ml_function_b200.models import keras_binary_crossentropy
    B, F, k = 65536, 26, 16
    torch.manual_seed(2020)
    model = bench.build_model("xdeepfm", DEV, cin_precision="bf16", mlp_dtype=torch.bfloat16)    # full Criteo tables
    # widen the (tiny, glorot) tables so that the interaction terms matter in the output
    with torch.no_grad():
        model.sparse_embed.arena.uniform_(-0.3, 0.3, generator=torch.Generator(device=DEV).manual_seed(1))
        model.linear_embed.arena.mul_(6.0)
    dense, ids, y = _batches(1, B, bench.CRITEO_ROWS, True, seed=9)[0]
    cap = {}
    front0 = model.front

    def front(d_, i_):
        ids_, xcat, v = front0(d_, i_)
        xcat.retain_grad()
        cap["xcat"] = xcat
        return ids_, xcat, v
    model.front = front
    out = model(dense, ids)                                           # [B,1,1]
    loss = keras_binary_crossentropy(y.view(out.shape), out)
    loss.backward()
    xcat = cap["xcat"]
    S = 48
    idx = torch.randperm(B, generator=gen(3))[:S].to(DEV)
    X = xcat.detach()[idx].double().cpu().requires_grad_(True)        # [S,W]: fields | dense | pad, physical order
    lin = model.linear_embed.lookup_sum(ids).detach()[idx].double().cpu().view(S, 1, 1)
    W = X.shape[1]
    E = X[:, :F * k].reshape(S, F, k)
    cw = [w.detach().double().cpu() for w in model.cin.conv_kernels]
    cb = [b.detach().double().cpu() for b in model.cin.conv_biases]
    cin_out = ko.cin(E, cw, cb, model.cin.logit_kernel.detach().double().cpu(), model.cin.logit_bias.detach().double().cpu())
    # The MLP runs in bf16: a hidden pre-activation within bf16 rounding distance of 0 flips its ReLU -- a property of
    # the kink, not of the kernels -- and one flipped unit moves that sample's input gradient by O(weight).  As for the
    # bf16 attention (test_fullsize_gpu / test_ref_pinned_gpu), the fp64 oracle therefore uses the ReLU masks of the
    # bf16 forward (recomputed here with the same GEMMs on the full batch); everything else is DnnLayer.call, CL:201-226.
    masks = []
    with torch.no_grad():
        h = xcat.detach().to(torch.bfloat16)
        for w, b in zip(model.dnn.kernels, model.dnn.biases):
            h = torch._addmm_activation(b.to(torch.bfloat16), h, w.to(torch.bfloat16))
            masks.append((h[idx] > 0).double().cpu())
    hx = X
    for w, b, mk in zip(model.dnn.kernels, model.dnn.biases, masks):
        hx = ko.keras_dense(hx, w.detach().double().cpu(), b.detach().double().cpu()) * mk
    dnn_out = ko.keras_dense(hx, model.dnn.logit_kernel.detach().double().cpu(), model.dnn.logit_bias.detach().double().cpu())
    dnn_plain = ko.dnn_layer(X.detach(), [w.detach().double().cpu() for w in model.dnn.kernels],
                             [b.detach().double().cpu() for b in model.dnn.biases],
                             model.dnn.logit_kernel.detach().double().cpu(), model.dnn.logit_bias.detach().double().cpu())
    assert (dnn_out.detach() - dnn_plain).abs().max() < 2e-2 * dnn_plain.abs().max()   # masks barely move the VALUE
    ref = ko.score_layer([lin, cin_out, dnn_out], use_add=True)       # MD:136
    assert ref.shape == (S, 1, 1)
    assert float(ref.max() - ref.min()) > 0.1 and 0.02 < float(ref.min()) and float(ref.max()) < 0.98   # not trivial / saturated
    assert_rel(out.detach()[idx], ref, 2e-2, "xdeepfm bench-config outputs (sampled rows)")
    lg = torch.log(ref / (1 - ref))
    assert_rel(model.logit(dense, ids).detach()[idx], lg, 2e-2, "xdeepfm bench-config logits (sampled rows)")
    # per-sample input gradient: d(mean_B bce)/d xcat[b] depends on sample b alone
    yl = y[idx].double().cpu().view(S, 1, 1)
    p = torch.clamp(ref, 1e-7, 1 - 1e-7)
    part = (-(yl * torch.log(p + 1e-7) + (1 - yl) * torch.log(1 - p + 1e-7))).sum() / B
    part.backward()
    e_emb = rel_err(xcat.grad[idx][:, :F * k], X.grad[:, :F * k])
    e_den = rel_err(xcat.grad[idx][:, F * k:F * k + 13], X.grad[:, F * k:F * k + 13])
    print(f"bench-config xDeepFM: out err {rel_err(out.detach()[idx], ref):.3e}, d/d(emb rows) {e_emb:.3e}, d/d(dense) {e_den:.3e}")
    assert e_emb <= 2e-2, f"xdeepfm bench-config d(loss)/d(embedding rows): rel err {e_emb:.3e} > 2e-2"
    assert e_den <= 2e-2, f"xdeepfm bench-config d(loss)/d(dense features): rel err {e_den:.3e} > 2e-2"
    del model
    torch.cuda.empty_cache()


def test_autoint_three_layer_bf16_stack_vs_stacked_oracle():
    from ml_function_b200 import layers as KL, models as KM
    from ml_function_b200.models import keras_binary_crossentropy
    g = gen(21)
    rows, k, H, dd, NL, B = SMALL_ROWS, 16, 2, 8, 3, 96
    p = {}
    for f, r in enumerate(rows):
        p[f"emb_{f}"] = torch.randn(r, k, generator=g) * 0.6
        p[f"lin_{f}"] = torch.randn(r, 1, generator=g) * 0.1
    for l in range(NL):
        sfx = "" if l == 0 else f"_{l}"
        for w in ("query_w", "key_w", "res_w"):
            p[w + sfx] = torch.randn(k, H, dd, generator=g) * 0.3
        p["ln_gamma" + sfx] = 1 + 0.2 * torch.randn(dd, generator=g)
        p["ln_beta" + sfx] = 0.1 * torch.randn(dd, generator=g)
    p["head_w"], p["head_b"] = ko.glorot_uniform((H * 26 * dd, 2), g), torch.randn(2, generator=g) * 0.1
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
    dense = torch.rand(B, 13, generator=g)
    yv = (torch.rand(B, generator=g) < 0.3).float()
    labels = torch.stack([1 - yv, yv], 1)
    p64 = {n: v.double().requires_grad_(True) for n, v in p.items()}
    out64 = ko.model_autoint_stacked(p64, dense.double(), ids, NL)
    ko.binary_crossentropy(labels.double(), out64).backward()
    sp = [KL.make_sparse_fea(str(14 + i), r, cross_unit=k) for i, r in enumerate(rows)]
    de = [KL.denseFea(str(1 + i), None) for i in range(13)]
    fea = KM.FeatureInput(sp, de, useLinear=True, device=DEV)
    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        model = KM.AutoInt(fea, attention_dim=dd, attention_head_dim=H, n_layers=NL, precision=precision)
        model.load_reference_params({n: v.to(DEV) for n, v in p.items()})
        out = model(dense.to(DEV), ids.to(DEV))
        assert_rel(out, out64, tol, f"autoint x{NL} {precision} forward")
        if precision == "fp32":     # gradients through three ReLU layers: compared where no kink can flip (fp32)
            keras_binary_crossentropy(labels.to(DEV), out).backward()
            got = model.reference_grads()
            for n in ("query_w", "key_w_1", "res_w_2", "ln_gamma_1", "head_w"):
                assert_rel(got[n], p64[n].grad, 2e-5, f"autoint x{NL} d{n}")
            offs = [0]
            for r in rows:
                offs.append(offs[-1] + r)
            dg = model.sparse_embed.arena.kon_sparse_grads[0].to_dense(offs[-1])
            assert_rel(dg, torch.cat([p64[f"emb_{f}"].grad for f in range(26)]), 2e-5, "autoint embedding grad")
            model.sparse_embed.arena.kon_sparse_grads = []
