"""Checkpoint / resume of a CTR model whose embedding tables are sharded over the ranks (the
reference's CTR examples never save; ``ModelCheckpoint(save_weights_only)`` appears only in its graph-embedding
trainer).

Layout of a checkpoint directory::

    manifest.json            world size, table rows, embedding dims, which rank owned which field and how
    dense.pt                 rank 0: every dense parameter (reference names where the model provides them),
                             the dense Adam state and the step counters
    emb_rank{r}.pt           rank r: its fields' tables -- whole tables for table-wise fields, the rows r::world for
                             row-wise fields -- and the row-wise Adam moments of the same rows

Every rank writes only what it owns (a 100 M-row table is never gathered).  ``load`` works for ANY target world size
and shard plan: each rank assembles exactly the rows it owns under the new plan from the shard files that hold them
(table-wise -> one file; row-wise -> interleaved from all old ranks), field by field, memory-mapped.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import torch


def _plan_of(embed, world: int):
    """(fields this rank holds, {field: ('tw'|'rw')}, arena offsets per held field)."""
    if hasattr(embed, "plan"):                               # parallel.ShardedEmbed
        plan = embed.plan
        fields = plan.tw_of_rank[embed.rank] + plan.rw_fields
        kind = {f: ("rw" if f in plan.rw_fields else "tw") for f in fields}
        offs = list(embed.all_offs)
        return fields, kind, offs
    F = len(embed.field_row_offset) - 1                      # layers.SparseEmbed: every table, whole
    return list(range(F)), {f: "tw" for f in range(F)}, list(embed.field_row_offset)


def _sparse_opt_of(trainer, arena):
    if trainer is None:
        return None
    for so in trainer.sparse_opts:
        if so.arena is arena:
            return so
    return None


def save(path: str, model, trainer=None, rank: int = 0, world: int = 1) -> None:
    os.makedirs(path, exist_ok=True)
    shard: Dict[str, dict] = {}
    owners: Dict[str, dict] = {}
    for which in ("sparse_embed", "linear_embed"):
        emb = getattr(model, which, None)
        if emb is None:
            continue
        fields, kind, offs = _plan_of(emb, world)
        so = _sparse_opt_of(trainer, emb.arena)
        ent = {}
        for j, f in enumerate(fields):
            sl = slice(offs[j], offs[j + 1])
            rec = {"kind": kind[f], "w": emb.arena.detach()[sl].cpu().clone()}
            if so is not None:
                rec["m"], rec["v"] = so.m[sl].cpu().clone(), so.v[sl].cpu().clone()
            ent[f] = rec
        shard[which] = {"fields": ent, "t": None if so is None else int(so.t)}
        owners[which] = {str(f): kind[f] for f in fields}
    torch.save(shard, os.path.join(path, f"emb_rank{rank}.pt"))
    # who holds what (every rank contributes its line; rank 0 merges after the barrier below)
    with open(os.path.join(path, f"owners_rank{rank}.json"), "w") as fh:
        json.dump(owners, fh)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0:
        sp = {id(p) for p in model.sparse_parameters()}
        dense = {n: p.detach().cpu() for n, p in model.named_parameters() if id(p) not in sp}
        state = {"dense": dense}
        if trainer is not None and trainer.dense_opt is not None:
            state["dense_opt"] = trainer.dense_opt.state_dict()
        torch.save(state, os.path.join(path, "dense.pt"))
        merged: Dict[str, Dict[str, list]] = {}
        for r in range(world):
            with open(os.path.join(path, f"owners_rank{r}.json")) as fh:
                for which, d in json.load(fh).items():
                    for f, kd in d.items():
                        merged.setdefault(which, {}).setdefault(f, []).append([r, kd])
        rows = [int(i.word_size) for i in model.sparse_embed.sparse_info] if hasattr(model.sparse_embed, "sparse_info") \
            else list(model.sparse_embed.plan.rows)
        with open(os.path.join(path, "manifest.json"), "w") as fh:
            json.dump({"world": world, "rows": rows, "owners": merged, "format": 1}, fh)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _field_table(path: str, manifest: dict, which: str, f: int, key: str, cache: dict) -> Optional[torch.Tensor]:
    """The full (global) tensor `key` ('w'|'m'|'v') of field f, assembled from the shard files."""
    holders = manifest["owners"][which][str(f)]
    world_old = manifest["world"]

    def shard_of(r):
        if r not in cache:
            cache[r] = torch.load(os.path.join(path, f"emb_rank{r}.pt"), map_location="cpu", mmap=True, weights_only=False)
        return cache[r][which]["fields"][f]
    if holders[0][1] == "tw":
        rec = shard_of(holders[0][0])
        return rec.get(key)
    parts = {r: shard_of(r).get(key) for r, _ in holders}
    if any(p is None for p in parts.values()):
        return None
    n = manifest["rows"][f]
    out = torch.empty((n,) + tuple(next(iter(parts.values())).shape[1:]), dtype=next(iter(parts.values())).dtype)
    for r, p in parts.items():
        out[r::world_old] = p
    return out


def load(path: str, model, trainer=None, rank: int = 0, world: int = 1) -> None:
    with open(os.path.join(path, "manifest.json")) as fh:
        manifest = json.load(fh)
    cache: dict = {}
    for which in ("sparse_embed", "linear_embed"):
        emb = getattr(model, which, None)
        if emb is None or which not in manifest["owners"]:
            continue
        fields, kind, offs = _plan_of(emb, world)
        so = _sparse_opt_of(trainer, emb.arena)
        for j, f in enumerate(fields):
            sl = slice(offs[j], offs[j + 1])
            for key, dst in (("w", emb.arena.data), ("m", None if so is None else so.m), ("v", None if so is None else so.v)):
                if dst is None:
                    continue
                t = _field_table(path, manifest, which, f, key, cache)
                if t is None:
                    continue
                if kind[f] == "rw":
                    t = t[rank::world]
                with torch.no_grad():
                    dst[sl].copy_(t.to(dst.device))
        if so is not None:
            t_saved = torch.load(os.path.join(path, "emb_rank0.pt"), map_location="cpu", mmap=True, weights_only=False)[which]["t"]
            if t_saved is not None:
                so.t.fill_(int(t_saved))
    state = torch.load(os.path.join(path, "dense.pt"), map_location="cpu", weights_only=False)
    own = dict(model.named_parameters())
    with torch.no_grad():
        for n, v in state["dense"].items():
            if n in own:
                own[n].data.copy_(v.to(own[n].device))
    if trainer is not None and "dense_opt" in state and trainer.dense_opt is not None:
        trainer.dense_opt.load_state_dict(state["dense_opt"])
