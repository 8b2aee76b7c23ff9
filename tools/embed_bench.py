"""Times the embedding kernels alone (Criteo cardinalities, B=65536, k=16) with an L2 flush between
iterations: kon_embed_fwd, and the pieces of kon_embed_bwd via the library's per-kernel events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_function_b200 import _lib as L, ops
from bench import CRITEO_ROWS

dev = "cuda:0"
B, k = 65536, 16
zipf = len(sys.argv) > 1 and sys.argv[1] == "zipf"
g = torch.Generator().manual_seed(0)
offs = [0]
for r in CRITEO_ROWS:
    offs.append(offs[-1] + r)
arena = torch.randn(offs[-1], k, device=dev)
cols = []
for r in CRITEO_ROWS:
    if zipf:
        u = torch.rand(B, generator=g, dtype=torch.float64)
        c = (torch.exp(u * torch.log(torch.tensor(float(r) + 1.0))) - 1.0).long().clamp_(0, r - 1)
    else:
        c = torch.randint(0, r, (B,), generator=g)
    cols.append(c)
ids = torch.stack(cols, 1).to(torch.int32).to(dev)
xcat = torch.empty(B, 432, device=dev)
out = xcat[:, :416].view(B, 26, k)
gout = torch.randn(B, 432, device=dev)[:, :416].view(B, 26, k)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = L.lib()
res = {}
for name, fn in (("fwd", lambda: ops.embed_fwd_raw(arena, ids, offs, out=out)),
                 ("bwd", lambda: ops.embed_bwd_raw(gout, ids, offs))):
    ts = []
    lib.kon_profile_reset(); lib.kon_profile_enable(1)
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    lib.kon_profile_enable(0)
    ts = sorted(ts[2:])
    res[name] = ts[len(ts) // 2]
    for kn in ("embed_fwd_vec_kernel", "embed_bwd_sort", "embed_reduce_kernel"):
        ms, n = L.profile_read(kn)
        if n:
            res[name + ":" + kn] = ms / n
# back to back from a CUDA graph (no host gaps): the whole backward, and the presorted backward (reduce + fixup only)
def graph_time(fn, reps=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(reps):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); gr.replay(); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res["bwd_graph"] = graph_time(lambda: ops.embed_bwd_raw(gout, ids, offs, share_sort=False))
ws = torch.empty(lib.kon_embed_bwd_workspace_bytes(ids.numel(), 32), dtype=torch.uint8, device=dev)
def sort_only():
    a, w = L._arg(ids), L._arg(ws)
    L.check(lib.kon_embed_sort(a.ptr, L.i64_array(offs), 26, w.ptr, torch.cuda.current_stream().cuda_stream), "sort")
res["sort_graph"] = graph_time(sort_only)
fwd_bytes = B * 26 * (4 + 2 * k * 4)
uniq = int(ops.embed_bwd_raw(gout, ids, offs).n.item())
bwd_bytes = B * 26 * (4 + k * 4) + uniq * (k * 4 + 4)
print(json.dumps({"ids": "zipf" if zipf else "uniform", "ms": {a: round(b, 4) for a, b in res.items()},
                  "fwd_GBs": fwd_bytes / res["fwd:embed_fwd_vec_kernel"] / 1e6,
                  "bwd_reduce_GBs": bwd_bytes / res["bwd:embed_reduce_kernel"] / 1e6,
                  "bwd_total_GBs": bwd_bytes / res["bwd"] / 1e6,
                  "bwd_graph_GBs": bwd_bytes / res["bwd_graph"] / 1e6, "unique_rows": uniq}))
