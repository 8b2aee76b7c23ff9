"""-m "not gpu": the C-ABI library loads and exports every symbol include/kon_b200.h declares,
rejects CPU tensors loudly (no fallback), and the host-side layer logic."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from ml_function_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return _lib


def test_header_symbols_all_exported():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "kon_b200.h")).read()
    declared = set(re.findall(r"\b(kon_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in kon_b200.h but not exported"
    assert declared == set(L.EXPORTED_SYMBOLS)
    assert lib.kon_abi_version() == 1


def test_ops_surface_is_complete():
    """Every op entry the layers / trainer / sharding code calls exists (a GPU-only code path must not be
    the first place a missing name shows up)."""
    from ml_function_b200 import ops
    for name in ("embed_fwd_raw", "embed_bwd_raw", "embed_fwd_peer", "embed_bwd_peer", "embed_lookup",
                 "embed_lookup_concat", "embed_sgd", "embed_adam", "embed_adam_devstep", "fm", "cross", "cin",
                 "attention", "new_step", "end_step", "profile_summary", "SparseGrad"):
        assert callable(getattr(ops, name)), name


def test_cpu_tensors_are_rejected_not_computed():
    L = _lib()
    from ml_function_b200 import ops
    v = torch.randn(4, 3, 8)
    with pytest.raises(L.KonError, match="not a CUDA tensor"):
        ops.fm(v, None)
    with pytest.raises(L.KonError):
        ops.embed_fwd_raw(torch.randn(10, 4), torch.zeros(2, 1, dtype=torch.int32), [0, 10])
    with pytest.raises(L.KonError):
        ops.cross(torch.randn(4, 8), torch.randn(2, 8), torch.randn(2, 8))


def test_missing_library_fails_loudly(monkeypatch):
    L = _lib()
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libkon_b200.so")
    with pytest.raises(L.KonError, match="no CPU or PyTorch fallback"):
        L.lib()


def test_workspace_queries_need_no_gpu():
    L = _lib()
    lib = L.lib()
    assert lib.kon_embed_bwd_workspace_bytes(65536 * 26, 16) > 65536 * 26 * 20
    hs = L.i32_array([200, 200, 200])
    assert lib.kon_cin_saved_bytes(128, 26, 16, hs, 3, L.KON_CIN_FP32) == 128 * 16 * 600 * 4


def test_pack_ids_reference_call_convention():
    from ml_function_b200.layers import pack_ids
    cols = [torch.tensor([[1.0], [2.0], [3.0]]), torch.tensor([[0.0], [5.0], [7.0]])]   # float32 Keras Inputs (DP:290)
    ids = pack_ids(cols)
    assert ids.dtype == torch.int32 and ids.shape == (3, 2) and ids[2, 1] == 7
    seq = [torch.zeros(3, 5, dtype=torch.int64), torch.ones(3, 5, dtype=torch.int64)]
    assert pack_ids(seq).shape == (3, 2, 5)


def test_sparse_fea_fields_match_reference():
    from ml_function_b200.layers import make_sparse_fea, sparseFea
    assert sparseFea._fields == ('fea_name', 'word_size', 'input_dim', 'cross_unit', 'linear_unit', 'pre_weight',
                                 'mask_zero', 'is_trainable', 'input_length', 'sample_num', 'batch_size', 'emb_reg')
    f = make_sparse_fea("14", 100)
    assert f.cross_unit == 8 and f.linear_unit == 1 and f.emb_reg == 1e-8 and f.input_length == 1


def test_ref_to_phys_row_permutation():
    from ml_function_b200 import layers as KL, models as KM
    sp = [KL.make_sparse_fea(str(i), 5, cross_unit=4) for i in range(3)]
    de = [KL.denseFea(str(i), None) for i in range(13)]
    fea = KM.FeatureInput(sp, de, useLinear=True, device="cpu")
    m = KM.DeepFM(fea, hidden_units=[8, 4])
    assert m.W == 28 and m.Fk == 12
    w = torch.arange(25, dtype=torch.float32).view(25, 1)
    ph = m.ref_to_phys_rows(w)
    assert ph.shape == (28, 1)
    assert ph[:12, 0].tolist() == list(range(13, 25)) and ph[12:25, 0].tolist() == list(range(13))
    assert ph[25:].abs().sum() == 0


def test_bench_reference_arm_line():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "fm",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_shard_plan_criteo_8_ranks():
    """8xB200 plan for the Criteo cardinalities: contiguous table-wise blocks (3-4 fields per rank), no
    permutation after the exchange, 100 M-row tables go row-wise (config 5)."""
    from ml_function_b200.parallel import ShardPlan
    from bench import CRITEO_ROWS
    plan = ShardPlan(CRITEO_ROWS, 8)
    assert plan.rw_fields == [] and plan.identity_order
    sizes = [len(fs) for fs in plan.tw_of_rank]
    assert sum(sizes) == 26 and max(sizes) - min(sizes) <= 1
    flat = [f for fs in plan.tw_of_rank for f in fs]
    assert flat == list(range(26))                       # contiguous blocks in field order
    big = ShardPlan(CRITEO_ROWS + [100_000_000, 100_000_000], 8)
    assert big.rw_fields == [26, 27] and big.identity_order
    assert sum(big.local_rows(r, 26) for r in range(8)) == 100_000_000
    mixed = ShardPlan([10, 60_000_000, 20], 4)
    assert mixed.rw_fields == [1] and not mixed.identity_order and mixed.exchange_order == [0, 2, 1]
    assert [mixed.to_global[f] for f in range(3)] == [0, 2, 1]


def test_routing_pass_plan_covers_every_id_bit_once():
    """kon_embed_route_plan (host logic of the embedding backward's per-field counting sort, no GPU): the digits of a
    table's passes tile its id bits exactly, no pass has more than 4096 digits, the last pass only has the digits the
    row count can produce, and a table's passes sit in the first slots of the job."""
    import ctypes
    from ml_function_b200 import _lib as L
    lib = L.lib()
    out = (ctypes.c_int64 * 6)()
    for rows in [0, 1, 2, 3, 4096, 4097, 65536, 10131227, (1 << 24), (1 << 24) + 1, 100_000_000, (1 << 31) - 2]:
        P = lib.kon_embed_route_plan(rows, 0, 3, out)
        bits = max(rows - 1, 0).bit_length() if rows > 1 else 0
        assert P == (1 if bits <= 12 else -(-bits // 12)), rows
        covered = 0
        for slot in range(3):
            lib.kon_embed_route_plan(rows, slot, 3, out)
            active, first, last, shift, bins, mask = (int(v) for v in out)
            assert active == (slot < P), (rows, slot)
            if not active:
                continue
            assert first == (slot == 0) and last == (slot == P - 1)
            assert 1 <= bins <= 4096
            assert shift == covered
            if last:
                assert bins == ((max(rows - 1, 0) >> shift) + 1)          # only digits that can occur
                covered = bits
            else:
                assert bins == mask + 1 and bins & (bins - 1) == 0
                covered += bins.bit_length() - 1
            # every id below `rows` lands in a digit of the pass
            for idv in {0, max(rows - 1, 0), max(rows // 2, 0)}:
                d = (idv >> shift) & mask
                assert 0 <= d < bins
        assert covered == bits, rows
