"""NVLink exchange micro-benchmark: what the sharded-embedding exchange achieves on the wire.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_bench.py [--out gpurun_out/r02_peer_bench_N.json]

One process per GPU, one ``PeerRegion`` each (CUDA-IPC mapped by every other rank).  Measured with CUDA events on
the launching stream, max over ranks:

* ``barrier``      -- ``kon_peer_barrier`` alone (flag stores + spin over NVLink), per call;
* ``put_big``      -- ``kon_peer_put2d`` of one contiguous 64 MiB block to every other rank (the store bandwidth the
                      exchange can reach at best);
* ``put_rows``     -- the training step's shape: ``B_l = 65,536`` rows, this rank's share of the 26 x 64-byte
                      embedding columns stored into the pitched ``[N*B_l, cols]`` gradient receive buffers of the owners
                      (backward) -- short rows, pitched destination;
* ``gather_push``  -- ``kon_embed_fwd_peer_cols``: the forward gather whose stores land in the sample owners' ``xcat``.

``nvidia-smi nvlink -gt d`` counters of GPU 0 are read before and after (rank 0): the delta shows the bytes really
travelled over NVLink and not through host memory."""
import argparse
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_function_b200 import _lib as L, ops                      # noqa: E402
from ml_function_b200.parallel import PeerRegion, _put2d          # noqa: E402


def nvlink_kib(idx=0):
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(idx)], capture_output=True, text=True,
                             timeout=20).stdout
    except Exception:
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return {"tx_kib": tx, "rx_kib": rx} if (tx or rx) else None


class _Sh:          # what _put2d needs
    def __init__(self, t):
        self.arena = t


def timed(fn, reg, iters, dev, group):
    for _ in range(3):
        fn()
    reg.barrier()
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    return float(ms.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    N, B_l, F, k = world, 65536, 26, 16
    big = 64 << 20
    # region: flags | big receive (N blocks) | pitched gradient receive [N*B_l, cols_max*k] | xcat [B_l, F*k+16]
    cols = [len(range(q, F, N)) for q in range(N)]                       # count-balanced field ownership
    W = F * k + 16
    nbytes = 256 + N * big + N * B_l * max(cols) * k * 4 + B_l * W * 4 + (1 << 20)
    reg = PeerRegion(group, dev, nbytes)
    recv_big, off_big = reg.carve((N, big // 4))
    drecv, off_d = reg.carve((N * B_l, max(cols) * k))                   # same carve on every rank; pitch = the owner's columns
    xcat, off_x = reg.carve((B_l, W))
    sh = _Sh(xcat)
    src_big = torch.randn(big // 4, device=dev)
    gout = torch.randn(B_l, W, device=dev)
    res = {"n_gpus": N, "B_l": B_l}
    nv0 = nvlink_kib(0) if rank == 0 else None

    res["barrier_us"] = 1e3 * timed(lambda: reg.barrier(), reg, 200, dev, group)

    def put_big():
        _put2d(sh, [(src_big.data_ptr(), reg.ptrs[q] + off_big + rank * big, big, big, big, 1)
                    for q in range(N) if q != rank])
        reg.barrier()
    ms = timed(put_big, reg, a.iters, dev, group)
    sent = (N - 1) * big
    res["put_big"] = {"ms": ms, "bytes_out_per_gpu": sent, "GBps_out_per_gpu": sent / ms / 1e6,
                      "note": "includes one barrier per iteration"}

    # backward shape: owner q's columns of my [B_l, W] gradient -> q's [N*B_l, cols_q*k] buffer, rows rank*B_l..
    col0 = [sum(cols[:q]) for q in range(N)]                              # fields laid out owner-major for the bench

    def put_rows():
        puts = []
        for q in range(N):
            if q == rank:
                continue
            wq = cols[q] * k * 4
            # every rank carved the same way, but the pitch of q's buffer is q's column count
            puts.append((gout.data_ptr() + col0[q] * k * 4, reg.ptrs[q] + off_d + rank * B_l * wq, W * 4, wq, wq, B_l))
        _put2d(sh, puts)
        reg.barrier()
    ms = timed(put_rows, reg, a.iters, dev, group)
    sent = sum(cols[q] * k * 4 * B_l for q in range(N) if q != rank)
    t = torch.tensor([float(sent)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    res["put_rows"] = {"ms": ms, "bytes_out_per_gpu_max": int(t.item()), "GBps_out_per_gpu": t.item() / ms / 1e6,
                       "row_bytes": [cols[q] * k * 4 for q in range(N)], "note": "includes one barrier per iteration"}

    # forward shape: gather my tables for the GLOBAL batch and store the rows into the owners' xcat
    rows_per_field = 1_000_000
    Fl = cols[rank]
    arena = torch.randn(Fl * rows_per_field, k, device=dev)
    ids = torch.randint(0, rows_per_field, (N * B_l, Fl), device=dev, dtype=torch.int32)
    offs = [f * rows_per_field for f in range(Fl + 1)]
    field_col = [(col0[rank] + f) * k for f in range(Fl)]
    peer_out = reg.ptr_array(off_x)

    def gather_push():
        ops.embed_fwd_peer(arena, ids, offs, peer_out, N, B_l, W, k, field_col=field_col)
        reg.barrier()
    ms = timed(gather_push, reg, a.iters, dev, group)
    sent = (N - 1) * B_l * Fl * k * 4
    t = torch.tensor([float(sent)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    res["gather_push"] = {"ms": ms, "bytes_out_per_gpu_max": int(t.item()), "GBps_out_per_gpu": t.item() / ms / 1e6,
                          "gathered_bytes_per_gpu": N * B_l * Fl * k * 4, "note": "includes one barrier per iteration"}
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)
    if rank == 0:
        nv1 = nvlink_kib(0)
        if nv0 and nv1:
            res["nvlink_gpu0_delta_MiB"] = {"tx": (nv1["tx_kib"] - nv0["tx_kib"]) / 1024, "rx": (nv1["rx_kib"] - nv0["rx_kib"]) / 1024}
            it = a.iters + 3
            res["expected_tx_MiB_gpu0"] = it * (res["put_big"]["bytes_out_per_gpu"] + sum(cols[q] * k * 4 * B_l for q in range(1, N))
                                                + (N - 1) * B_l * cols[0] * k * 4) / 2 ** 20
        else:
            res["nvlink_gpu0_delta_MiB"] = None
        line = json.dumps(res)
        print(line)
        if a.out:
            with open(a.out, "w") as f:
                f.write(line + "\n")
    reg.check()
    reg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
