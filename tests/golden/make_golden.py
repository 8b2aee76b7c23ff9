"""Generates tests/golden/kon_golden.npz: small seeded input/weight/output vectors of every
hot-path layer, computed by the CPU oracle (oracle/kon_oracle.py) in the reference's op order
(fp32) and by the fp64 closed forms.

The reference ships no golden vectors and TensorFlow cannot be imported here, so
these pin the ORACLE (against silent drift) and give the -m gpu tests a fixed target that does
not depend on the oracle being importable; they are not outputs of the reference itself
("parity unpinned").  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kon_oracle as ko  # noqa: E402


def main():
    g = torch.Generator().manual_seed(2020)
    out = {}
    # embeddings: 3 fields, dim 8
    rows = [7, 50, 3]
    B, k = 33, 8
    tables = [torch.randn(r, k, generator=g) for r in rows]
    lins = [torch.randn(r, 1, generator=g) for r in rows]
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
    idl = [ids[:, f:f + 1] for f in range(3)]
    emb = ko.sparse_embed(idl, tables, use_flatten=False)
    lin = ko.sparse_embed(idl, lins, use_flatten=False)
    out["emb_tables"] = torch.cat(tables).numpy()
    out["emb_lins"] = torch.cat(lins).numpy()
    out["emb_rows"] = np.array(rows)
    out["emb_ids"] = ids.numpy()
    out["emb_out"] = torch.cat(emb, 1).numpy()
    d_out = torch.randn(B, 3, k, generator=g)
    out["emb_dout"] = d_out.numpy()
    offs = np.concatenate([[0], np.cumsum(rows)])
    ur, ug = [], []
    for f in range(3):
        u, gr = ko.embedding_grad(ids[:, f].numpy(), d_out[:, f].numpy(), rows[f])
        ur.append(u + offs[f]); ug.append(gr)
    out["emb_unique_rows"] = np.concatenate(ur).astype(np.int32)
    out["emb_grads"] = np.concatenate(ug)
    # FM
    out["fm_out"] = ko.fm_layer(emb, lin)[:, 0].numpy()
    out["fm_out_f64"] = ko.fm_closed_form(torch.cat(emb, 1).double(), torch.cat(lin, 1)[..., 0].double()).numpy()
    # Cross
    D, Lc = 29, 3
    x = torch.randn(B, D, generator=g)
    cw = [torch.randn(D, 1, generator=g) * 0.3 for _ in range(Lc)]
    cb = [torch.randn(D, 1, generator=g) * 0.1 for _ in range(Lc)]
    out["cross_x"], out["cross_w"], out["cross_b"] = x.numpy(), torch.stack(cw).numpy(), torch.stack(cb).numpy()
    out["cross_out"] = ko.cross_layer(x, cw, cb).numpy()
    # CIN
    m, Dk, hs = 5, 4, [6, 7]
    x0 = torch.randn(B, m, Dk, generator=g)
    hp, ws, bs = m, [], []
    for n in hs:
        ws.append(torch.randn(1, hp * m, n, generator=g) * 0.2)
        bs.append(torch.randn(n, generator=g) * 0.1)
        hp = n
    out["cin_x0"] = x0.numpy()
    for i in range(2):
        out[f"cin_w{i}"], out[f"cin_b{i}"] = ws[i].numpy(), bs[i].numpy()
    out["cin_pooled"] = ko.cin(x0, ws, bs, return_pooled=True).numpy()
    out["cin_pooled_f64"] = ko.cin_closed_form(x0.double(), [w.double() for w in ws], [b.double() for b in bs])[0].numpy()
    # attention block
    F, kin, H, d = 6, 8, 2, 4
    xa = torch.randn(B, F, kin, generator=g)
    wq, wk, wr = (torch.randn(kin, H, d, generator=g) * 0.4 for _ in range(3))
    gam, bet = torch.rand(d, generator=g) + 0.5, torch.randn(d, generator=g) * 0.1
    out["attn_x"], out["attn_wq"], out["attn_wk"], out["attn_wr"] = xa.numpy(), wq.numpy(), wk.numpy(), wr.numpy()
    out["attn_gamma"], out["attn_beta"] = gam.numpy(), bet.numpy()
    out["attn_out"] = ko.autoint_block(xa, wq, wk, wr, gam, bet).numpy()
    # 2-unit head over two inputs (MergeScoreLayer, CL:86-100) and NFM end to end (MD:108-119); appended after
    # the original draws, so every earlier array keeps its bits
    x1, x2 = torch.randn(B, 12, generator=g), torch.randn(B, 5, generator=g)
    hw, hb = torch.randn(17, 2, generator=g) * 0.3, torch.randn(2, generator=g) * 0.1
    out["head_x1"], out["head_x2"], out["head_w"], out["head_b"] = x1.numpy(), x2.numpy(), hw.numpy(), hb.numpy()
    out["head_logits"] = ko.keras_dense(torch.cat([x1, x2], 1), hw, hb).numpy()
    out["head_softmax"] = ko.merge_score_layer([x1, x2], hw, hb).numpy()
    p = {}
    for f in range(3):
        p[f"emb_{f}"], p[f"lin_{f}"] = tables[f], lins[f]
    dims = [13 + k, 10, 6, 4]
    for i in range(3):
        p[f"dnn_w{i}"] = torch.randn(dims[i], dims[i + 1], generator=g) * 0.3
        p[f"dnn_b{i}"] = torch.randn(dims[i + 1], generator=g) * 0.1
    p["dnn_logit_w"], p["dnn_logit_b"] = torch.randn(4, 1, generator=g) * 0.5, torch.zeros(1)
    dense = torch.rand(B, 13, generator=g)
    out["nfm_dense"] = dense.numpy()
    for i in range(3):
        out[f"nfm_dnn_w{i}"], out[f"nfm_dnn_b{i}"] = p[f"dnn_w{i}"].numpy(), p[f"dnn_b{i}"].numpy()
    out["nfm_logit_w"] = p["dnn_logit_w"].numpy()
    out["nfm_out"] = ko.model_nfm(p, dense, ids).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kon_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
