"""-m "not gpu": the multi-rank routing of the sharded embedding (world_size 2, gloo, CPU).

The exchange logic (`parallel.py`) is exercised with torch-CPU stand-ins for the two kernels it
calls (`kon_embed_fwd` / `kon_embed_bwd`) -- test infrastructure only; the product path binds the
CUDA kernels and has no CPU implementation."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import offsets


def _cpu_lookup(arena, ids, offs, sum_fields=False):
    B, F = ids.shape
    out = torch.zeros(B, F, arena.shape[1])
    for f in range(F):
        rows = offs[f + 1] - offs[f]
        i = ids[:, f].long()
        ok = (i >= 0) & (i < rows)
        out[ok, f] = arena[(offs[f] + i[ok])]
    return out.sum(1) if sum_fields else out


def _cpu_scatter(g, ids, offs):
    from ml_function_b200.ops import SparseGrad
    B, F = ids.shape
    dense = torch.zeros(offs[-1], g.shape[-1])
    for f in range(F):
        rows = offs[f + 1] - offs[f]
        i = ids[:, f].long()
        ok = (i >= 0) & (i < rows)
        dense.index_add_(0, offs[f] + i[ok], g[ok, f])
    nz = dense.abs().sum(1).nonzero().flatten()
    return SparseGrad(nz.int(), dense[nz], torch.tensor([nz.numel()], dtype=torch.int32))


def _worker(rank, world, port, rows, k, B_l, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_function_b200 import layers as KL
    from ml_function_b200.parallel import ShardPlan, ShardedEmbed
    g = torch.Generator().manual_seed(1)
    tables = [torch.randn(r, k, generator=g) for r in rows]
    lins = [torch.randn(r, 1, generator=g) for r in rows]
    ids_all = torch.stack([torch.randint(0, r, (world * B_l,), generator=g) for r in rows], 1).to(torch.int32)
    gout_all = torch.randn(world * B_l, len(rows), k, generator=g)
    info = [KL.make_sparse_fea(str(i), r, cross_unit=k) for i, r in enumerate(rows)]
    plan = ShardPlan(rows, world, row_wise_min_rows=100)
    sh = ShardedEmbed(info, plan, dist.group.WORLD, "cpu", lookup_fn=_cpu_lookup, scatter_fn=_cpu_scatter)
    sh.load_global_tables(tables)
    ids = ids_all[rank * B_l:(rank + 1) * B_l]
    emb = sh.lookup(ids)
    ref = torch.stack([tables[f][ids[:, f].long()] for f in range(len(rows))], 1)
    ok_fwd = torch.equal(emb, ref)
    emb.backward(gout_all[rank * B_l:(rank + 1) * B_l])
    # expected: this rank's shard of the GLOBAL gradient
    dense = torch.zeros(sh.arena.shape)
    fields = plan.tw_of_rank[rank] + plan.rw_fields
    for j, f in enumerate(fields):
        full = torch.zeros(rows[f], k).index_add_(0, ids_all[:, f].long(), gout_all[:, f])
        if f in plan.rw_fields:
            full = full[rank::world]
        dense[sh.all_offs[j]:sh.all_offs[j + 1]] = full
    got = torch.zeros(sh.arena.shape)
    for sg in sh.arena.kon_sparse_grads:
        got.index_add_(0, sg.rows[:int(sg.n)].long(), sg.grads[:int(sg.n)])
    ok_bwd = torch.allclose(got, dense, atol=1e-5)
    # first-order tables: sum over fields
    shl = ShardedEmbed(info, plan, dist.group.WORLD, "cpu", is_linear=True, lookup_fn=_cpu_lookup, scatter_fn=_cpu_scatter)
    shl.load_global_tables(lins)
    s = shl.lookup_sum(ids)
    ref_s = sum(lins[f][ids[:, f].long()] for f in range(len(rows)))
    ok_sum = torch.allclose(s, ref_s, atol=1e-5)
    ret[rank] = (ok_fwd, ok_bwd, ok_sum, len(plan.rw_fields), len(plan.tw_fields))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_embedding_routing_gloo(world):
    rows = [7, 300, 5, 1000, 2, 150, 33]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + world + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, rows, 4, 6, ret), nprocs=world, join=True)
    for r in range(world):
        ok_fwd, ok_bwd, ok_sum, n_rw, n_tw = ret[r]
        assert ok_fwd and ok_bwd and ok_sum, (r, ret[r])
        assert n_rw == 3 and n_tw == 4


def test_shard_plan_covers_every_row_once():
    from ml_function_b200.parallel import ShardPlan
    rows = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
            14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
    for world in (1, 2, 4, 8):
        for thr in (50_000_000, 1_000_000):
            p = ShardPlan(rows, world, row_wise_min_rows=thr)
            assert sorted(p.exchange_order) == list(range(26))
            for f in range(26):
                assert sum(p.local_rows(r, f) for r in range(world) if f in p.rw_fields or p.tw_owner[f] == r) == rows[f]
            cnt = [len(x) for x in p.tw_of_rank]
            assert max(cnt) - min(cnt) <= 1                      # balanced by lookup count
            if thr == 50_000_000:                                # Criteo: everything table-wise, contiguous blocks
                assert p.rw_fields == [] and p.identity_order
            elif world == 8:
                assert len(p.rw_fields) == 5


def test_cost_balanced_plan_spreads_the_big_tables():
    from ml_function_b200.parallel import ShardPlan
    rows = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
            14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
    big = [f for f, r in enumerate(rows) if r >= 1_000_000]
    for world in (2, 4, 8):
        p = ShardPlan(rows, world, balance="cost")
        assert sorted(f for fs in p.tw_of_rank for f in fs) == list(range(26))          # every table exactly once
        per_rank_big = [sum(1 for f in fs if f in big) for fs in p.tw_of_rank]
        assert max(per_rank_big) - min(per_rank_big) <= 1                               # count plan: rank 0 had 2, four ranks 0
        assert max(len(fs) for fs in p.tw_of_rank) <= (26 + world - 1) // world         # NVLink egress stays capped
        for fs in p.tw_of_rank:
            assert fs == sorted(fs)
            rr = ShardPlan.runs(fs)
            assert sum(c for _, _, c in rr) == len(fs) and [f for f0, _, c in rr for f in range(f0, f0 + c)] == fs
            assert [j for _, j, _ in rr] == [fs.index(f0) for f0, _, _ in rr]
    assert ShardPlan(rows, 8).tw_of_rank[0] == [0, 1, 2, 3]                             # default stays contiguous


def _worker_step_cache(rank, world, port, ret):
    """Inside a training step (ops.new_step() ... end_step()) the embedding and the first-order tables share
    ONE ids exchange; outside a step nothing is cached (a recycled address must never hit)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_function_b200 import layers as KL, ops
    from ml_function_b200 import parallel as P
    rows = [7, 30, 5, 41]
    info = [KL.make_sparse_fea(str(i), r, cross_unit=4) for i, r in enumerate(rows)]
    plan = P.ShardPlan(rows, world)
    emb = P.ShardedEmbed(info, plan, dist.group.WORLD, torch.device("cpu"), lookup_fn=None, scatter_fn=None)
    lin = P.ShardedEmbed(info, plan, dist.group.WORLD, torch.device("cpu"), is_linear=True, lookup_fn=None, scatter_fn=None)
    calls = {"n": 0}
    real = P._all_to_all

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    P._all_to_all = counting
    ids = torch.stack([torch.randint(0, r, (6,)) for r in rows], 1).to(torch.int32)
    a = emb.exchange_ids(ids)
    b = lin.exchange_ids(ids)
    outside = calls["n"]
    ops.new_step()
    c = emb.exchange_ids(ids)
    d = lin.exchange_ids(ids)
    inside = calls["n"] - outside
    ops.end_step()
    e = emb.exchange_ids(ids)
    after = calls["n"] - outside - inside
    ret[rank] = (outside, inside, after, c[0] is d[0], torch.equal(a[0], c[0]) and torch.equal(a[0], e[0]), emb.use_peer)
    dist.destroy_process_group()


def test_ids_exchange_is_shared_within_a_step_only():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_step_cache, args=(2, 29931 + (os.getpid() % 300), ret), nprocs=2, join=True)
    for r in range(2):
        outside, inside, after, same_obj, same_val, use_peer = ret[r]
        assert outside == 2 and inside == 1 and after == 1 and same_obj and same_val and not use_peer
