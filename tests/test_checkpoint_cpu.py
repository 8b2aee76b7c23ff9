"""-m "not gpu": checkpoint of sharded tables (gloo, world 2) and re-sharding on load (world 2 -> 1 and 2 -> 3)."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROWS = [7, 300, 5, 1000, 2, 150, 33]
K = 4


class _Opt:
    def __init__(self, arena):
        self.arena = arena
        g = torch.Generator().manual_seed(int(arena.numel()))
        self.m, self.v = torch.randn(arena.shape, generator=g), torch.rand(arena.shape, generator=g)
        self.t = torch.tensor([7], dtype=torch.int32)


class _Trainer:
    def __init__(self, model):
        self.sparse_opts = [_Opt(model.sparse_embed.arena), _Opt(model.linear_embed.arena)]
        self.dense_opt = None


class _Model(nn.Module):
    def __init__(self, sparse, linear):
        super().__init__()
        self.sparse_embed, self.linear_embed = sparse, linear
        self.w = nn.Parameter(torch.arange(6.0).reshape(2, 3))

    def sparse_parameters(self):
        return [self.sparse_embed.arena, self.linear_embed.arena]


def _tables(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(r, K, generator=g) for r in ROWS], [torch.randn(r, 1, generator=g) for r in ROWS]


def _build(world, rank, group):
    from ml_function_b200 import layers as KL
    from ml_function_b200.parallel import ShardPlan, ShardedEmbed
    info = [KL.make_sparse_fea(str(i), r, cross_unit=K) for i, r in enumerate(ROWS)]
    plan = ShardPlan(ROWS, world, row_wise_min_rows=100)
    return _Model(ShardedEmbed(info, plan, group, "cpu", lookup_fn=None, scatter_fn=None),
                  ShardedEmbed(info, plan, group, "cpu", is_linear=True, lookup_fn=None, scatter_fn=None))


def _global(model, trainer, rank, world, which, key):
    """{field: this rank's view of the global table} for comparison."""
    emb = getattr(model, which)
    plan = emb.plan
    src = {"w": emb.arena.data, "m": trainer.sparse_opts[0 if which == "sparse_embed" else 1].m,
           "v": trainer.sparse_opts[0 if which == "sparse_embed" else 1].v}[key]
    out = {}
    for j, f in enumerate(plan.tw_of_rank[rank] + plan.rw_fields):
        out[f] = (src[emb.all_offs[j]:emb.all_offs[j + 1]].clone(), f in plan.rw_fields)
    return out


def _worker(rank, world, port, path, mode, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_function_b200 import checkpoint
    model = _build(world, rank, dist.group.WORLD)
    tr = _Trainer(model)
    tabs, lins = _tables(3)
    if mode == "save":
        model.sparse_embed.load_global_tables(tabs)
        model.linear_embed.load_global_tables(lins)
        checkpoint.save(path, model, tr, rank=rank, world=world)
        ret[rank] = True
    else:
        checkpoint.load(path, model, tr, rank=rank, world=world)
        ok = True
        for which, ref in (("sparse_embed", tabs), ("linear_embed", lins)):
            for f, (t, is_rw) in _global(model, tr, rank, world, which, "w").items():
                want = ref[f][rank::world] if is_rw else ref[f]
                ok = ok and torch.equal(t, want)
        ok = ok and int(tr.sparse_opts[0].t) == 7 and torch.equal(model.w.data, torch.arange(6.0).reshape(2, 3) + 1)
        ret[rank] = ok
    dist.destroy_process_group()


def _run(world, path, mode):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29700 + (os.getpid() % 200) + world, path, mode, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_save_world2_then_load_into_world2_world3_and_single():
    from ml_function_b200 import checkpoint
    with tempfile.TemporaryDirectory() as d:
        # make the dense parameter distinguishable from its init
        assert all(_run(2, d, "save"))
        st = torch.load(os.path.join(d, "dense.pt"), weights_only=False)
        st["dense"]["w"] += 1
        torch.save(st, os.path.join(d, "dense.pt"))
        assert sorted(os.listdir(d)) == ["dense.pt", "emb_rank0.pt", "emb_rank1.pt", "manifest.json", "owners_rank0.json",
                                         "owners_rank1.json"]
        assert all(_run(2, d, "load"))            # same plan
        assert all(_run(3, d, "load"))            # re-sharded: 3 row-wise shards from 2
        # single process, unsharded SparseEmbed-like layout
        tabs, lins = _tables(3)

        class _Flat(nn.Module):
            def __init__(self, ts):
                super().__init__()
                offs = [0]
                for t in ts:
                    offs.append(offs[-1] + t.shape[0])
                self.field_row_offset = tuple(offs)
                self.arena = nn.Parameter(torch.zeros(offs[-1], ts[0].shape[1]))
        m = _Model(_Flat(tabs), _Flat(lins))
        tr = _Trainer(m)
        checkpoint.load(d, m, tr, rank=0, world=1)
        assert torch.equal(m.sparse_embed.arena.data, torch.cat(tabs)) and torch.equal(m.linear_embed.arena.data, torch.cat(lins))
        # Adam moments travelled with the rows: compare against what rank-level generators produced at save time
        assert int(tr.sparse_opts[1].t) == 7
