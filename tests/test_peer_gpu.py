"""-m gpu, one GPU: the peer-exchange entry points (kon_embed_fwd_peer / kon_embed_bwd_peer /
kon_peer_barrier) with the "peers" being slabs in this GPU's own memory -- the addressing,
the skip-invalid rule of row-wise shards and the barrier protocol are the same; the NVLink mappings
themselves are covered by tests/test_parallel_gpu.py on >= 2 GPUs."""
import ctypes

import pytest
import torch

from helpers import gen, make_ids, make_tables, offsets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ptrs(tensors, byte_off=0):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() + byte_off for t in tensors])


@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
@pytest.mark.parametrize("N,B_l,dim", [(2, 96, 16), (4, 33, 8), (8, 1000, 16), (3, 50, 32)])
def test_embed_fwd_peer_matches_local(N, B_l, dim, idt):
    from ml_function_b200 import ops
    rows = [7, 300, 5, 41]
    g = gen(5)
    arena = torch.cat(make_tables(rows, dim, g)).to(DEV)
    offs = offsets(rows)
    ids = make_ids(N * B_l, rows, g, dtype=idt).to(DEV)
    ref = ops.embed_fwd_raw(arena, ids, offs)                          # [N*B_l, 4, dim]
    # each "rank" holds a [B_l, W] concat buffer; this rank's 4 fields start at column 2*dim
    W = 7 * dim + 4
    bufs = [torch.full((B_l, W), -7.0, device=DEV) for _ in range(N)]
    ops.embed_fwd_peer(arena, ids, offs, _ptrs(bufs, 2 * dim * 4), N, B_l, W, dim)
    torch.cuda.synchronize()
    for q in range(N):
        got = bufs[q][:, 2 * dim:6 * dim].view(B_l, 4, dim)
        assert torch.equal(got, ref[q * B_l:(q + 1) * B_l]), q
        assert (bufs[q][:, :2 * dim] == -7.0).all() and (bufs[q][:, 6 * dim:] == -7.0).all()


def test_embed_fwd_peer_skip_invalid_is_row_wise_assembly():
    """Row-wise shards: every rank gathers with the rows of other ranks marked -1 and
    KON_EMBED_SKIP_INVALID; the union of the N launches assembles the full lookup."""
    from ml_function_b200 import ops
    N, B_l, dim, R = 4, 64, 16, 1001
    g = gen(6)
    table = make_tables([R], dim, g)[0].to(DEV)
    ids = make_ids(N * B_l, [R, R], g).to(DEV)                           # two row-wise fields, same table size
    bufs = [torch.full((B_l, 2 * dim), float("nan"), device=DEV) for _ in range(N)]
    for r in range(N):
        shard = table[r::N].contiguous()
        arena = torch.cat([shard, shard])
        lo = [0, shard.shape[0], 2 * shard.shape[0]]
        loc = torch.where(ids % N == r, ids // N, torch.full_like(ids, -1))
        ops.embed_fwd_peer(arena, loc.contiguous(), lo, _ptrs(bufs), N, B_l, 2 * dim, dim, skip_invalid=True)
    torch.cuda.synchronize()
    got = torch.cat(bufs).view(N * B_l, 2, dim)
    assert torch.equal(got, table[ids.long()])


@pytest.mark.parametrize("N,B_l,dim", [(2, 96, 16), (4, 257, 8), (8, 500, 16)])
def test_embed_bwd_peer_matches_local(N, B_l, dim):
    from ml_function_b200 import ops
    rows = [3, 900, 17, 1, 50000]
    F = len(rows)
    g = gen(7)
    offs = offsets(rows)
    ids = make_ids(N * B_l, rows, g).to(DEV)
    ids[::5, 1] = -1                                                     # rows of "another rank": no gradient
    W = (F + 3) * dim
    gbufs = [torch.randn(B_l, W, generator=g).to(DEV) for _ in range(N)]
    col0 = 2 * dim
    d_out = torch.cat([b[:, col0:col0 + F * dim] for b in gbufs]).view(N * B_l, F, dim).contiguous()
    ref = ops.embed_bwd_raw(d_out, ids, offs, share_sort=False)
    got = ops.embed_bwd_peer(_ptrs(gbufs, col0 * 4), N, B_l, W, dim, dim, ids, offs)
    torch.cuda.synchronize()
    n = int(ref.n.item())
    assert int(got.n.item()) == n
    assert torch.equal(got.rows[:n], ref.rows[:n])
    assert torch.equal(got.grads[:n], ref.grads[:n])                     # same order of summation: bit-exact


def test_peer_entry_points_reject_bad_arguments():
    from ml_function_b200 import _lib as L, ops
    arena = torch.randn(10, 6, device=DEV)                               # dim 6: no 128-bit rows
    ids = torch.zeros(4, 1, dtype=torch.int32, device=DEV)
    buf = torch.zeros(4, 8, device=DEV)
    with pytest.raises(L.KonError, match="dim % 4"):
        ops.embed_fwd_peer(arena, ids, [0, 10], _ptrs([buf]), 1, 4, 8, 8)
    arena = torch.randn(10, 8, device=DEV)
    with pytest.raises(L.KonError, match="peers hold"):
        ops.embed_fwd_peer(arena, ids, [0, 10], _ptrs([buf]), 1, 2, 8, 8)
    with pytest.raises(L.KonError, match="n_peers"):
        ops.embed_fwd_peer(arena, ids, [0, 10], _ptrs([buf] * 17), 17, 4, 8, 8)


def _barrier(lib, L, flags, rank, stream, timeout_ms=2000):
    L.check(lib.kon_peer_barrier(_ptrs(flags), len(flags), rank, 0, timeout_ms, stream.cuda_stream), "kon_peer_barrier")


def test_peer_barrier_two_ranks_on_two_streams_and_timeout():
    from ml_function_b200 import _lib as L
    lib = L.lib()
    flags = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(2)]
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(5):                       # five epochs; the epoch counter lives on the device
        _barrier(lib, L, flags, 0, s0)
        _barrier(lib, L, flags, 1, s1)
    torch.cuda.synchronize()
    for r in range(2):
        assert flags[r][16].item() == 5 and flags[r][17].item() == 0
        assert flags[r][0].item() == 5 and flags[r][1].item() == 5
    # rank 1 never arrives: rank 0 gives up after the timeout and reports it, the GPU is not hung
    _barrier(lib, L, flags, 0, s0, timeout_ms=50)
    torch.cuda.synchronize()
    assert flags[0][17].item() == 2          # 1 + index of the missing peer


def test_peer_alloc_handle_free():
    from ml_function_b200 import _lib as L
    lib = L.lib()
    ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
    L.check(lib.kon_peer_alloc(0, 1 << 20, ctypes.byref(ptr), handle), "kon_peer_alloc")
    assert ptr.value and any(bytes(handle))
    from ml_function_b200.parallel import _RawCuda
    t = torch.as_tensor(_RawCuda(ptr.value, 1 << 20), device=DEV)
    assert t.data_ptr() == ptr.value and int(t.sum().item()) == 0          # zero-initialised, aliased not copied
    del t
    L.check(lib.kon_peer_free(0, ptr), "kon_peer_free")
