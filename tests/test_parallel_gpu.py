"""-m gpu (needs >= 2 GPUs, skipped otherwise): sharded model == single-GPU model, NCCL."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import gen

pytestmark = pytest.mark.gpu


def _build(name, rows, k, dev):
    from ml_function_b200 import layers as KL, models as KM
    sp = [KL.make_sparse_fea(str(14 + i), r, cross_unit=k) for i, r in enumerate(rows)]
    de = [KL.denseFea(str(1 + i), None) for i in range(13)]
    fea = KM.FeatureInput(sp, de, useLinear=True, useAddLinear=(name == "xdeepfm"), device=dev)
    if name == "xdeepfm":
        return KM.XDeepFM(fea, conv_size=[16, 8], hidden_units=[32, 16], cin_precision="fp32")
    return KM.DeepFM(fea, hidden_units=[32, 16])


def _worker(rank, world, port, name, peer, ret):
    os.environ["KON_PEER_EXCHANGE"] = peer      # "1": fused NVLink exchange kernels, "0": NCCL all-to-all baseline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ml_function_b200.parallel import DistContext
    from ml_function_b200.train import Trainer
    rows = [7, 3000, 5, 41, 2, 1500, 26, 900]
    k, B_l = 8, 96
    g = gen(3)
    ids_all = torch.stack([torch.randint(0, r, (world * B_l,), generator=g) for r in rows], 1).to(torch.int32)
    dense_all = torch.rand(world * B_l, 13, generator=g)
    y = (torch.rand(world * B_l, generator=g) < 0.3).float()
    labels_all = y.view(-1, 1) if name == "xdeepfm" else torch.stack([1 - y, y], 1)
    # single-GPU reference model (same seeds -> same dense weights), global batch
    ref = _build(name, rows, k, dev)
    tables = [ref.sparse_embed.arena.detach()[ref.sparse_embed.field_row_offset[f]:ref.sparse_embed.field_row_offset[f + 1]].clone()
              for f in range(len(rows))]
    lins = [ref.linear_embed.arena.detach()[ref.linear_embed.field_row_offset[f]:ref.linear_embed.field_row_offset[f + 1]].clone()
            for f in range(len(rows))]
    for t in tables + lins:
        dist.broadcast(t, 0)
    ref.sparse_embed.load_reference_weights(tables)
    ref.linear_embed.load_reference_weights(lins)
    out_ref = ref(dense_all.to(dev), ids_all.to(dev))
    # sharded model, local batch
    model = _build(name, rows, k, dev)
    dctx = DistContext(dist.group.WORLD, dev, row_wise_min_rows=500)
    dctx.attach(model)
    model.sparse_embed.load_global_tables(tables)
    model.linear_embed.load_global_tables(lins)
    sl = slice(rank * B_l, (rank + 1) * B_l)
    out = model(dense_all[sl].to(dev), ids_all[sl].to(dev))
    err_fwd = (out - out_ref[sl]).abs().max().item()
    # one training step on both; compare a dense weight and the loss
    tr_ref, tr = Trainer(ref, lr=1e-2), Trainer(model, lr=1e-2, dist_ctx=dctx)
    for _ in range(2):      # two steps: the exchange buffers are reused
        l_ref = tr_ref.step(dense_all.to(dev), ids_all.to(dev), labels_all.to(dev))
        l_loc = tr.step(dense_all[sl].to(dev), ids_all[sl].to(dev), labels_all[sl].to(dev))
    lt = l_loc.clone()
    dist.all_reduce(lt)
    err_loss = abs(lt.item() / world - l_ref.item())
    err_w = (model.dnn.kernels[0].detach() - ref.dnn.kernels[0].detach()).abs().max().item()
    # embedding rows after the step: gather this rank's shard from the reference arena
    plan = dctx.plan
    fields = plan.tw_of_rank[rank] + plan.rw_fields
    err_e = 0.0
    for j, f in enumerate(fields):
        lo, hi = ref.sparse_embed.field_row_offset[f], ref.sparse_embed.field_row_offset[f + 1]
        full = ref.sparse_embed.arena.detach()[lo:hi]
        if f in plan.rw_fields:
            full = full[rank::world]
        mine = model.sparse_embed.arena.detach()[model.sparse_embed.all_offs[j]:model.sparse_embed.all_offs[j + 1]]
        err_e = max(err_e, (mine - full).abs().max().item())
    assert model.sparse_embed.use_peer == (peer == "1")
    if peer == "1":
        assert not any(px["region"].timed_out() for px in model.sparse_embed._peer.values())
        model.sparse_embed.close_peer()
    ret[rank] = (err_fwd, err_loss, err_w, err_e)
    dist.destroy_process_group()


@pytest.mark.parametrize("peer", ["1", "0"])
@pytest.mark.parametrize("name", ["deepfm", "xdeepfm"])
def test_sharded_model_matches_single_gpu(name, peer):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29611 + (os.getpid() % 500) + (7 if peer == "1" else 0), name, peer, ret), nprocs=world, join=True)
    for r in range(world):
        err_fwd, err_loss, err_w, err_e = ret[r]
        assert err_fwd < 1e-5 and err_loss < 1e-5 and err_w < 1e-4 and err_e < 1e-4, (r, ret[r])
