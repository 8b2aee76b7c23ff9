"""Per-kernel device time of the training step (torch.profiler / CUPTI), single GPU or under
torchrun.  Diagnostic only: numbers taken under a profiler are never bench values.

    python tools/step_profile.py [--model xdeepfm] [--steps 4]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_profile.py
"""
import argparse, os, sys
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from ml_function_b200.train import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="xdeepfm")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--rows", type=int, default=30)
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dctx = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    from ml_function_b200.parallel import DistContext
    dctx = DistContext(dist.group.WORLD, dev)
torch.manual_seed(2020)
model = bench.build_model(args.model, dev, mlp_dtype=torch.bfloat16)
if dctx is not None:
    dctx.attach(model)
tr = Trainer(model, lr=1e-3, dist_ctx=dctx)
host = bench.synth_batches(2, args.batch, bench.CRITEO_ROWS, 2020 + rank, args.model == "xdeepfm")
res = [tuple(t.to(dev) for t in b) for b in host]
for i in range(5):
    tr.step(*res[i % 2])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
import time
def sync():
    torch.cuda.synchronize()
    return time.perf_counter()
t0 = sync()
for i in range(args.steps):
    tr.step(*res[i % 2])
t1 = sync()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(args.steps):
        tr.step(*res[i % 2])
    torch.cuda.synchronize()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"step_profile_{args.model}_n{world}" + ("_nccl" if os.environ.get("KON_PEER_EXCHANGE") == "0" else "") + ".txt"), "w") as f:
        f.write(f"eager wall ms/step (no profiler): {(t1 - t0) / args.steps * 1e3:.3f}\n")
        f.write(prof.key_averages().table(sort_by="device_time_total", row_limit=args.rows, max_name_column_width=70))
if world > 1:
    if hasattr(model.sparse_embed, "close_peer"):
        model.sparse_embed.close_peer()
    dist.destroy_process_group()
