#!/usr/bin/env python
"""bench.py -- train-step throughput of the CTR hot path on synthetic Criteo-shaped data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model xdeepfm|deepfm|dcn|autoint|fm]
    python bench.py --impl reference ...      # the reference-equivalent CPU path (oracle port)
    python bench.py --gpus 8 --model autoint --big-tables 8x100000000     # BASELINE config 5 (row-sharded tables)
    python bench.py --gpus 8 --scaling strong                             # global batch 65,536 split over the ranks

Besides the headline (xDeepFM) line the default run adds ``other_models`` (DeepFM / DCN / AutoInt through the
same trainer, CUDA-graph replay, resident inputs) and, at N > 1, ``verified``: the sharded job against a
single-GPU model on the same global batch (forward, loss, dense-weight gradients, embedding-row gradients).

A "step" = one full training step (embedding gather -> interaction layers + MLP -> loss ->
backward incl. the sort-then-segment embedding scatter-add -> Adam on dense weights and
row-wise Adam on the touched embedding rows) over one batch of 65,536 synthetic samples
(26 sparse + 13 dense).  Metric: train samples/s (BASELINE.json), whole job over all ranks.

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM.  ``e2e``: the same step
through the public API with the batch coming from pinned host memory (H2D inside the timed
region) and the loss read back (D2H).  ``roofline``: the dominant kernel family, timed live
with CUDA events on the launching stream inside the timed region.  ``cpu_baseline``: the
oracle port (torch-CPU restatement of the reference layers; TensorFlow is not installable
here) on a bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os

# NCCL's version/debug banner goes to stdout by default; stdout carries exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CRITEO_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
               14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
N_DENSE = 13

MODEL_CFG = {
    # name: (emb dim, description)  -- BASELINE.json configs[1..4]
    "deepfm": dict(k=16, desc="DeepFM, 26 sparse + 13 dense, emb 16, MLP 256-128-64"),
    "dcn": dict(k=32, desc="DCN-v1, 6 CrossLayers, emb 32, MLP 256-128-64"),
    "xdeepfm": dict(k=16, desc="xDeepFM, CIN [200,200,200] bf16 tcgen05, emb 16, MLP 256-128-64"),
    "autoint": dict(k=16, desc="AutoInt, 3 layers x 2 heads x d=8 (bf16 tensor-core attention), emb 16"),
    "fm": dict(k=16, desc="FM, emb 16"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sust=1400.0, src="fallback")


# Tensor-pipe kernels are timed in windows of a fraction of a second at full boost clocks, so their roofline
# denominator is the BURST bf16 peak (the sustained figure is measured after seconds of back-to-back GEMMs at
# ~1350 MHz and would flatter the fraction); `frac_of_sustained` is reported beside it for reference.
TC_PEAK = "tc_burst"


# ------------------------------------------------------------------------------------------
# algorithmic bytes / flops per op per step (DESIGN.md "Roofline accounting")
# ------------------------------------------------------------------------------------------
def algo_work(model, B, k, conv=(200, 200, 200), cross_layers=6, heads=2, d=8, kin=16):
    F, p = 26, 4
    W = (F * k + N_DENSE + 3) // 4 * 4
    w = {
        "embed_fwd": ("hbm", B * F * (4 + 2 * k * p)),
        "embed_bwd": ("hbm", B * F * (4 + k * p) + B * F * (k * p + 4)),       # worst case all rows unique
        "fm_fwd": ("hbm", B * (F * k * p + F * p + k * p)),
        "fm_bwd": ("hbm", B * (k * p + 2 * F * k * p + F * p)),
        "cross_fwd": ("hbm", B * (2 * W * p + 4 * cross_layers)),
        "cross_bwd": ("hbm", B * 3 * W * p),
        "attn_fwd": ("hbm", B * (F * kin * p + heads * F * d * p)),
        "attn_bwd": ("hbm", B * (2 * F * kin * p + heads * F * d * p)),
    }
    hd = {"deepfm": k + 64, "fm": k, "dcn": W + 64, "autoint": heads * F * d}.get(model)
    if hd:                                                     # MergeScoreLayer's Dense(2) over hd columns
        w["head_fwd"] = ("hbm", B * (hd * p + 2 * p))
        w["head_bwd"] = ("hbm", B * (2 * hd * p + 2 * p))
    hp, fl, per_layer = F, 0, []
    for n in conv:
        per_layer.append(2 * B * k * hp * F * n)
        fl += per_layer[-1]
        hp = n
    # What the tcgen05 GEMM kernels actually execute.  The reference pools the LAST layer over its feature maps
    # (IL:322) and nothing consumes z_L, so the last layer needs no GEMM in either direction: forward
    # pool_L = (pre (x) x0) . rowsum(W) + sum(b) (cin_last_pool_kernel), backward dZ_L = broadcast gradient
    # (cin_last_dw_kernel / cin_last_da_kernel).  Only layers 0..L-2 run as GEMMs and only their FLOPs are
    # credited; the dense-GEMM definition of the whole CIN (fl forward, 2*fl backward) is NOT used for any fraction.
    gemm = sum(per_layer[:-1]) if len(per_layer) >= 2 else fl
    w["cin_fwd_gemm"] = ("tensor", gemm)
    w["cin_dw_gemm"] = ("tensor", gemm)
    w["cin_da_gemm"] = ("tensor", gemm)
    if len(conv) >= 2:      # last-layer forward: reads pre [B,H,k] bf16 + x0 [B,F,k] f32, writes pooled [B,k] f32
        w["cin_last_pool"] = ("hbm", B * (conv[-2] * k * 2 + F * k * p + k * p))
    return w


# library kernel (or kernel family bracketed by one event pair in the library) ->
# (op whose algorithmic work it carries, share of that op's work, work is per LAUNCH rather than per step)
KERNEL_WORK = {
    "cin_fwd_tc_kernel": ("cin_fwd_gemm", 1.0, False),    # layers 0..L-2 (last layer: pooled mat-vec below)
    "cin_last_pool_kernel": ("cin_last_pool", 1.0, False),
    "cin_dw_tc_kernel": ("cin_dw_gemm", 1.0, False),      # layers 0..L-2 (last layer: rank-1 shortcut)
    "cin_da_tc_kernel": ("cin_da_gemm", 1.0, False),
    "embed_fwd_vec_kernel": ("embed_fwd", 1.0, False),
    "embed_reduce_kernel": ("embed_bwd", 1.0, False),
    "fm_fwd_kernel": ("fm_fwd", 1.0, True),
    "fm_bwd_kernel": ("fm_bwd", 1.0, True),
    "cross_fwd_kernel": ("cross_fwd", 1.0, True),
    "cross_bwd_kernels": ("cross_bwd", 1.0, True),        # per-sample kernel + dw kernel + reduce + finalize
    "attn_tc_fwd_kernel": ("attn_fwd", 1.0, True),        # one launch per attention layer
    "attn_tc_bwd_kernel": ("attn_bwd", 1.0, True),
    "head_fwd_kernel": ("head_fwd", 1.0, True),
    "head_bwd_kernels": ("head_bwd", 1.0, True),
}


def sample_clocks_start(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                stdout=f, stderr=subprocess.DEVNULL), f
    except Exception:
        return None, None


def sample_clocks_stop(proc, f, path, dev_index):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        proc.wait(timeout=5)
    except Exception:
        proc.kill()
    f.close()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in open(path):
        c = [x.strip() for x in line.split(",")]
        if len(c) < 9 or not c[0].isdigit() or int(c[0]) != dev_index:
            continue
        try:
            sm.append(float(c[1])); mx.append(float(c[2]))
        except ValueError:
            continue
        for n, v in zip(names, c[5:9]):
            if v.lower().startswith("active"):
                reasons.add(n)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm.sort()
    # "under load": upper half of the samples (the sampler also sees the idle edges)
    load = sm[len(sm) // 2:]
    return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def nvlink_counters(dev_index):
    """Cumulative NVLink data bytes (tx, rx) of one GPU, summed over its links: `nvidia-smi nvlink -gt d`
    (the NVML throughput counters; KiB).  None when the tool / counters are not there."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(dev_index)], capture_output=True,
                             text=True, timeout=20).stdout
    except Exception:
        return None
    tx = rx = 0
    seen = False
    for line in out.splitlines():
        m = re.search(r"Data (Tx|Rx):\s*(\d+)\s*KiB", line)
        if m:
            seen = True
            if m.group(1) == "Tx":
                tx += int(m.group(2)) * 1024
            else:
                rx += int(m.group(2)) * 1024
    return (tx, rx) if seen else None


# ------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------
def synth_batches(n, B, rows, seed, sigmoid_head, zipf=False, pin=True):
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        cols = []
        for r in rows:
            if zipf:
                u = torch.rand(B, generator=g, dtype=torch.float64)
                c = (torch.exp(u * torch.log(torch.tensor(float(r) + 1.0))) - 1.0).long().clamp_(0, r - 1)
            else:
                c = torch.randint(0, r, (B,), generator=g)
            cols.append(c)
        ids = torch.stack(cols, 1).to(torch.int32).contiguous()
        dense = torch.rand(B, N_DENSE, generator=g)
        y = (torch.rand(B, generator=g) < 0.25).float()
        labels = y.view(B, 1) if sigmoid_head else torch.stack([1 - y, y], 1)       # to_categorical, DP:359
        t = (dense.contiguous(), ids, labels.contiguous())
        if pin:
            t = tuple(x.pin_memory() for x in t)
        out.append(t)
    return out


def build_model(name, device, cin_precision="bf16", rows=CRITEO_ROWS, mlp_dtype=None, lazy_tables=False):
    """``lazy_tables``: the embedding layers are about to be replaced by sharded ones (DistContext.attach), so the
    replicated arenas are created on the meta device (a 100M-row table must never be materialised unsharded)."""
    import torch
    from ml_function_b200 import layers as KL
    from ml_function_b200 import models as KM
    k = MODEL_CFG[name]["k"]
    sparse = [KL.make_sparse_fea(str(14 + i), r, cross_unit=k) for i, r in enumerate(rows)]
    dense = [KL.denseFea(str(1 + i), None) for i in range(N_DENSE)]
    fea = KM.FeatureInput(sparse, dense, useLinear=True, useAddLinear=(name == "xdeepfm"),
                          device="meta" if lazy_tables else device)
    if name == "deepfm":
        m = KM.DeepFM(fea)
    elif name == "dcn":
        m = KM.DCN(fea, cross_hidden=6)
    elif name == "xdeepfm":
        m = KM.XDeepFM(fea, cin_precision=cin_precision)
    elif name == "autoint":
        m = KM.AutoInt(fea, attention_dim=8, attention_head_dim=2, n_layers=3, precision="bf16")
    else:
        m = KM.FM(fea)
    if mlp_dtype is not None and hasattr(m, "dnn"):
        m.dnn.compute_dtype = mlp_dtype
    return m


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same model (reference-equivalent CPU path)
# ------------------------------------------------------------------------------------------
def oracle_params(name, rows, k, gen):
    import torch
    from oracle import kon_oracle as ko
    p = {}
    F = len(rows)
    for f, r in enumerate(rows):
        p[f"emb_{f}"] = ko.glorot_uniform((r, k), gen)
        p[f"lin_{f}"] = (torch.rand(r, 1, generator=gen) - 0.5) * 0.1
    D = N_DENSE + F * k
    dims = [D, 256, 128, 64]
    for i in range(3):
        p[f"dnn_w{i}"] = ko.glorot_uniform((dims[i], dims[i + 1]), gen)
        p[f"dnn_b{i}"] = ko.glorot_uniform((dims[i + 1],), gen)
    if name == "deepfm":
        p["head_w"], p["head_b"] = ko.glorot_uniform((k + 64, 2), gen), torch.zeros(2)
    elif name == "fm":
        p["head_w"], p["head_b"] = ko.glorot_uniform((k, 2), gen), torch.zeros(2)
    elif name == "dcn":
        for i in range(6):
            p[f"outer_weight_{i}"] = ko.glorot_uniform((D, 1), gen)
            p[f"outer_bias_{i}"] = torch.zeros(D, 1)
        p["head_w"], p["head_b"] = ko.glorot_uniform((D + 64, 2), gen), torch.zeros(2)
    elif name == "xdeepfm":
        hp = F
        for i, n in enumerate((200, 200, 200)):
            p[f"cin_w{i}"] = ko.glorot_uniform((1, hp * F, n), gen)
            p[f"cin_b{i}"] = torch.zeros(n)
            hp = n
        p["cin_logit_w"], p["cin_logit_b"] = ko.glorot_uniform((3 * k, 1), gen), torch.zeros(1)
        p["dnn_logit_w"], p["dnn_logit_b"] = ko.glorot_uniform((64, 1), gen), torch.zeros(1)
    elif name == "autoint":
        H, d = 2, 8
        for w in ("query_w", "key_w", "res_w"):
            p[w] = ko.glorot_uniform((k, H, d), gen)
        p["ln_gamma"], p["ln_beta"] = torch.ones(d), torch.zeros(d)
        p["head_w"], p["head_b"] = ko.glorot_uniform((H * F * d, 2), gen), torch.zeros(2)
    return p


def oracle_step(name, p, dense, ids, labels, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7, l2=1e-8):
    """fwd + bwd + the optimizer step the reference's ``compile(optimizer='adam')`` performs, in the reference's op
    order (oracle/kon_oracle.py): Keras Adam is NOT lazy -- with ``embeddings_regularizer=l2`` (IL:217) the table
    gradient is dense (``2*l2*W`` everywhere) and every row's m / v / weight is touched each step."""
    import torch
    from oracle import kon_oracle as ko
    fn = {"fm": ko.model_fm, "deepfm": ko.model_deepfm, "xdeepfm": ko.model_xdeepfm, "autoint": ko.model_autoint,
          "dcn": lambda p_, d_, i_: ko.model_dcn(p_, d_, i_, cross_hidden=6)}[name]
    for t in p.values():
        t.requires_grad_(True)
        t.grad = None
    out = fn(p, dense, ids)
    loss = ko.binary_crossentropy(labels.view(out.shape), out)
    loss.backward()
    state["t"] = state.get("t", 0) + 1
    t_ = state["t"]
    lr_t = lr * (1 - b2 ** t_) ** 0.5 / (1 - b1 ** t_)
    with torch.no_grad():
        for key, w in p.items():
            if w.grad is None:
                continue
            g = w.grad
            if key.startswith("emb_"):
                g = g.add(w, alpha=2 * l2)                   # the regulariser's gradient: dense
            if key not in state:
                state[key] = (torch.zeros_like(w), torch.zeros_like(w))
            m, v = state[key]
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            w.addcdiv_(m, v.sqrt().add_(eps), value=-lr_t)
    return float(loss.detach())


def cpu_arm(name, sample_B, steps, warmup, seed=2020, tables="auto"):
    """Times the oracle port on all host cores, with the reference's optimizer (dense, non-lazy Adam + the L2 term).
    Tables: the full Criteo cardinalities when the host has the memory for weights + dense gradients + Adam state
    (~5x the table bytes), else capped at 200k rows per field (stated in the returned note)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    k = MODEL_CFG[name]["k"]
    full_bytes = sum(CRITEO_ROWS) * (k + 1) * 4
    if tables == "auto":
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:      # noqa: BLE001
            avail = 0
        tables = "full" if avail > 6 * full_bytes + (8 << 30) else "capped"
    rows = list(CRITEO_ROWS) if tables == "full" else [min(r, 200_000) for r in CRITEO_ROWS]
    g = torch.Generator().manual_seed(seed)
    p = oracle_params(name, rows, k, g)
    batches = synth_batches(2, sample_B, rows, seed, name == "xdeepfm", pin=False)
    state = {}
    for i in range(warmup):
        d, ids, y = batches[i % 2]
        oracle_step(name, p, d, ids, y, state)
    t0 = time.perf_counter()
    for i in range(steps):
        d, ids, y = batches[i % 2]
        oracle_step(name, p, d, ids, y, state)
    dt = time.perf_counter() - t0
    note = ("dense non-lazy Adam + L2 on the tables, as Keras applies them; " +
            ("full Criteo tables (33.76M rows)" if tables == "full" else "tables capped at 200k rows/field (host memory)"))
    return sample_B * steps / dt, dt / steps, torch.get_num_threads(), note


CPU_SAMPLE_B = {"xdeepfm": 1024, "deepfm": 16384, "dcn": 16384, "autoint": 16384, "fm": 16384}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.model
    sb = CPU_SAMPLE_B[name]
    v, s_per_step, cores, note = cpu_arm(name, sb, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(name, args, max(args.gpus, 1)),
        "run": {"batch_per_step": sb, "optimizer": "dense non-lazy Adam + dense L2 term (what Keras does)", "dtype": "f32"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {sb} (oracle port of the reference layers -- pinned bit for "
                                   f"bit to the reference's own source -- reference op order, torch-CPU fp32, fwd+bwd+{note})"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def shared_config(name, args, world):
    """`config` of the JSON line: the WORKLOAD, identical for the GPU arm and the `--impl reference` arm (which
    processes a bounded sample of it per step, stated in its cpu_baseline.sample).  How each arm runs it
    (parallelism, optimizer flavour, dtypes, timing) goes into `run`."""
    strong = getattr(args, "scaling", "weak") == "strong"
    B = args.batch // world if strong else args.batch
    rows = table_rows(args)
    return {"workload": MODEL_CFG[name]["desc"], "batch_per_gpu": B, "global_batch": B * world,
            "emb_dim": MODEL_CFG[name]["k"],
            "tables": ("Criteo-Kaggle cardinalities (33.76M rows)" if not args.big_tables else
                       f"Criteo-Kaggle cardinalities with {args.big_tables} rows ({sum(rows) / 1e6:.1f}M rows)"),
            "ids": args.ids, "loss": "binary_crossentropy", "optimizer": "Adam, lr 1e-3, L2 1e-8 on the embedding tables"}


def table_rows(args):
    """Criteo cardinalities; ``--big-tables TxR`` replaces the T largest tables by R-row ones (BASELINE config 5:
    100 M-row tables, row-wise sharded over the ranks when N > 1)."""
    rows = list(CRITEO_ROWS)
    if args.big_tables:
        t, r = args.big_tables.lower().split("x")
        t, r = int(t), int(r)
        for f in sorted(range(len(rows)), key=lambda f_: -rows[f_])[:t]:
            rows[f] = r
    return rows


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


class Job:
    """One model + trainer on this rank (sharded when world > 1), with its synthetic batches."""

    def __init__(self, name, args, dev, world, rank, B, rows, mlp_dtype, row_wise_min_rows=50_000_000):
        import torch
        import torch.distributed as dist
        from ml_function_b200.train import Trainer
        self.name, self.dev, self.world, self.rank, self.B = name, dev, world, rank, B
        torch.manual_seed(2020)
        self.model = build_model(name, dev, cin_precision=args.cin_precision, rows=rows, mlp_dtype=mlp_dtype,
                                 lazy_tables=world > 1)
        self.dctx = None
        if world > 1:
            from ml_function_b200.parallel import DistContext
            self.dctx = DistContext(dist.group.WORLD, dev, row_wise_min_rows=row_wise_min_rows)
            self.dctx.attach(self.model)
        self.trainer = Trainer(self.model, lr=1e-3, dist_ctx=self.dctx)
        self.host = synth_batches(args.n_batches, B, rows, 2020 + rank, name == "xdeepfm", zipf=(args.ids == "zipf"))
        self.resident = [tuple(t.to(dev) for t in b) for b in self.host]

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY `steps` steps between two CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks."""
        import torch
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def spin_up(self, fn, seconds):
        """Untimed steps for ~`seconds` of device time right before the timed windows (on top of --warmup).  With
        N > 1 the ranks wait on each other while rank 0 verifies / captures, their GPUs fall back to idle clocks, and
        the first 0.2 s window of a max-over-ranks timing then measures the clock ramp (8.18 / 7.76 / 7.54 ms for three
        consecutive windows at N = 2, gpurun_out/r02q_bench_n2.json) instead of the steady state."""
        if seconds <= 0:
            return
        ms10 = self.timed(fn, 10)                    # max over ranks: every rank derives the same step count
        n = int(min(2000, max(0, seconds * 1e3 / max(ms10 / 10, 1e-3) - 10)))
        for i in range(n):
            fn(i)
        self.barrier()

    def step_eager(self, i):
        d, ids, y = self.resident[i % len(self.resident)]
        return self.trainer.step(d, ids, y)

    def step_graph(self, i):
        d, ids, y = self.resident[i % len(self.resident)]
        return self.trainer.step_graph(d, ids, y)

    def check_peer(self):
        emb = self.model.sparse_embed
        if hasattr(emb, "_peer") and any(px["region"].timed_out() for px in emb._peer.values()):
            raise SystemExit(f"bench.py: rank {self.rank}: a peer barrier timed out; the measurement is void")

    def close(self):
        import torch
        self.trainer.release_graph()
        torch.cuda.synchronize()
        for emb in (self.model.sparse_embed, self.model.linear_embed):
            if emb is not None and hasattr(emb, "close_peer"):
                emb.close_peer()
        self.model = self.trainer = self.resident = self.host = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def quick_measure(name, args, dev, world, rank, B, rows, mlp_dtype, windows=3, row_wise_min_rows=50_000_000):
    """Compact measurement of one model through the same trainer: resident inputs, CUDA-graph replay, the median
    of `windows` windows of args.steps steps.  Returns a small dict (rank 0 prints it inside ``other_models``)."""
    job = Job(name, args, dev, world, rank, B, rows, mlp_dtype, row_wise_min_rows)
    try:
        for i in range(max(args.warmup, 3)):
            job.step_eager(i)
        graphed = False
        step = job.step_eager
        if args.graph and (world == 1 or args.graph_multi):
            graphed = job.trainer.capture(*job.resident[0])
            if graphed:
                step = job.step_graph
                for i in range(3):
                    step(i)
        job.spin_up(step, args.spinup)
        ms = median([job.timed(step, args.steps) for _ in range(windows)])
        job.check_peer()
        out = {"value": B * world * args.steps / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms / args.steps,
               "n_gpus": world, "batch_per_gpu": B, "cuda_graph": graphed, "workload": MODEL_CFG[name]["desc"],
               "parallelism": "single GPU" if world == 1 else job.dctx.describe()}
        if not graphed and args.graph:
            out["capture_error"] = job.trainer.capture_error
        return out
    finally:
        job.close()


def verify_sharded(name, args, dev, world, rank, B, rows, mlp_dtype):
    """The sharded job (N ranks, local batches) against a SINGLE-GPU model with the same weights on the same
    global batch, one forward + backward: per-sample outputs of rank 0's batch, the global loss, every dense
    weight gradient after the all-reduce, and the embedding-row gradients of the tables rank 0 owns.  Runs inside
    the driver's own scaling run, so every SCALE line proves sharded == single (the >= 2-GPU pytest cases are
    skipped by a 1-GPU test box)."""
    import torch
    import torch.distributed as dist
    from ml_function_b200 import ops
    from ml_function_b200.models import keras_binary_crossentropy
    from ml_function_b200.parallel import DistContext
    torch.manual_seed(2020)
    ref = build_model(name, dev, cin_precision=args.cin_precision, rows=rows, mlp_dtype=mlp_dtype) if rank == 0 else None
    torch.manual_seed(2020)
    model = build_model(name, dev, cin_precision=args.cin_precision, rows=rows, mlp_dtype=mlp_dtype, lazy_tables=True)
    dctx = DistContext(dist.group.WORLD, dev)
    dctx.attach(model)
    F = len(rows)
    for which in ("sparse_embed", "linear_embed"):
        sh = getattr(model, which)
        if sh is None:
            continue
        src = getattr(ref, which) if rank == 0 else None
        for f in range(F):
            if rank == 0:
                t = src.arena.detach()[src.field_row_offset[f]:src.field_row_offset[f + 1]]
            else:
                t = torch.empty((rows[f], sh.dim), device=dev)
            dist.broadcast(t, 0)
            sh.load_global_table(f, t)
            del t
    sig = name == "xdeepfm"
    mine = synth_batches(1, B, rows, 2020 + rank, sig, zipf=(args.ids == "zipf"), pin=False)[0]
    d, ids, y = (t.to(dev) for t in mine)
    ops.new_step(presort=False)
    out = model(d, ids)
    loss = keras_binary_crossentropy(y.view(out.shape), out)
    (loss / world).backward()
    dparams = model.dense_parameters()
    dctx.allreduce_dense_grads(dparams)
    ops.end_step()
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    res = None
    if rank == 0:
        rp = None
        losses, out0 = [], None
        for r in range(world):
            bd, bi, by = (t.to(dev) for t in synth_batches(1, B, rows, 2020 + r, sig, zipf=(args.ids == "zipf"), pin=False)[0])
            ops.new_step(presort=False)
            o = ref(bd, bi)
            lr_ = keras_binary_crossentropy(by.view(o.shape), o)
            (lr_ / world).backward()
            ops.end_step()
            losses.append(float(lr_.detach()))
            if r == 0:
                out0 = o.detach()
        rp = ref.dense_parameters()

        def rel(a, b):
            den = float(b.abs().max())
            return float((a - b).abs().max()) / (den if den > 0 else 1.0)
        e_fwd = rel(out.detach(), out0)
        e_loss = abs(float(lt) / world - sum(losses) / world)
        e_w = max(rel(p.grad, q.grad) for p, q in zip(dparams, rp) if q.grad is not None)
        # embedding rows: rank 0's shard of the sparse gradient vs the single-GPU one (summed over the N chunks)
        plan = dctx.plan
        sh, se = model.sparse_embed, ref.sparse_embed
        n_loc = sh.all_offs[-1]
        loc = sum(sg.to_dense(n_loc) for sg in sh.arena.kon_sparse_grads)
        e_e = 0.0
        touched_mismatch = 0
        fields = plan.tw_of_rank[0] + plan.rw_fields
        for j, f in enumerate(fields):
            lo, hi = se.field_row_offset[f], se.field_row_offset[f + 1]
            full = torch.zeros((hi - lo, se.dim), device=dev)
            for sg in se.arena.kon_sparse_grads:
                n = int(sg.n.item())
                r_ = sg.rows[:n].long()
                sel = (r_ >= lo) & (r_ < hi)
                full.index_add_(0, r_[sel] - lo, sg.grads[:n][sel])
            if f in plan.rw_fields:
                full = full[0::world]
            mine_f = loc[sh.all_offs[j]:sh.all_offs[j + 1]]
            e_f = rel(mine_f, full)
            if os.environ.get("KON_VERIFY_DETAIL"):
                dd = (mine_f - full).abs().amax(dim=1)
                bad = dd > 1e-5 * float(full.abs().max())
                print(f"verify detail: field {f} rows {rows[f]} err {e_f:.3e} max|ref| {float(full.abs().max()):.3e} "
                      f"bad rows {int(bad.sum())} of {int((full.abs().amax(dim=1) > 0).sum())} touched; "
                      f"touched-mismatch {int(((mine_f.abs().amax(dim=1) > 0) != (full.abs().amax(dim=1) > 0)).sum())}", file=sys.stderr)
            e_e = max(e_e, e_f)
            touched_mismatch += int(((mine_f.abs().amax(dim=1) > 0) != (full.abs().amax(dim=1) > 0)).sum())
            del full
        res = {"fwd": e_fwd, "loss": e_loss, "dense_w": e_w, "emb_rows": e_e, "emb_rows_touched_mismatch": touched_mismatch,
               # Routing is exact: the SET of touched rows must be identical (touched_mismatch == 0).  The values agree
               # to ~1e-3 of a field's largest entry, not to fp32 rounding: the first-order sum is associated
               # differently (per-owner partials), which moves the logit by an ulp, and in the bf16 CIN backward an ulp
               # flips the bf16 rounding of a dZ element now and then (observed: 0.4 % of the rows, |diff| ~ 2e-9).
               "ok": bool(e_fwd < 1e-5 and e_loss < 1e-5 and e_w < 2e-3 and e_e < 5e-3 and touched_mismatch == 0),
               "what": "max rel err, sharded N-rank job vs single-GPU model on the same global batch: rank-0 outputs, "
                       "global loss, dense-weight grads after all-reduce (summation order over the batch differs), "
                       "rank-0-owned embedding-row grads"}
    for emb in (model.sparse_embed, model.linear_embed):
        if emb is not None and hasattr(emb, "close_peer"):
            emb.close_peer()
    del model, ref
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    dist.barrier()
    return res


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ml_function_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("KON_BENCH_WATCHDOG"):       # debugging aid: dump every thread's stack and exit after N s
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["KON_BENCH_WATCHDOG"]), exit=True, file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    name = args.model
    strong = args.scaling == "strong"
    B = args.batch // world if strong else args.batch
    k = MODEL_CFG[name]["k"]
    rows = table_rows(args)
    mlp_dtype = {"bf16": torch.bfloat16, "f32": None, "tf32": None}[args.mlp_dtype]
    if args.mlp_dtype == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True

    # ---- N > 1: sharded == single GPU, before anything is timed ---------------------------------
    verified = None
    if world > 1 and args.verify and not args.big_tables:
        try:
            verified = verify_sharded(name, args, dev, world, rank, B, rows, mlp_dtype)
        except Exception as e:      # noqa: BLE001 -- never lose the measurement to the checker
            verified = {"ok": False, "error": repr(e)}
            print("verify_sharded failed:", repr(e), file=sys.stderr)
            torch.cuda.synchronize()

    rw_min = 50_000_000
    job = Job(name, args, dev, world, rank, B, rows, mlp_dtype, rw_min)
    model, trainer, host, resident, dctx = job.model, job.trainer, job.host, job.resident, job.dctx
    timed, step_resident, step_resident_graph = job.timed, job.step_eager, job.step_graph

    stage = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    # end to end: every step moves its batch host -> device (pinned memory) and reads the loss back.  The copies are
    # double-buffered on a copy stream, like the reference's dataset.prefetch(2) (DP:335-337): while step i runs, the
    # batch of step i + 1 is already on its way; one batch is copied per step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    staged = [None, None]          # which host batch sits (or is arriving) in stage[s]

    def prefetch(i):
        s = i % 2
        copy_stream.wait_event(ev_free[s])               # the step that last read stage[s] is done with it
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(stage[s], host[i % len(host)]):
                dst.copy_(src, non_blocking=True)
            ev_copied[s].record(copy_stream)
        staged[s] = i % len(host)

    def step_e2e(i):
        s = i % 2
        if staged[s] != i % len(host):
            prefetch(i)
        if args.e2e_prefetch:
            prefetch(i + 1)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev_copied[s])
        loss = trainer.step_graph(*stage[s])
        ev_free[s].record(cur)
        staged[s] = None                                  # consumed
        loss_host.copy_(loss.view(1), non_blocking=True)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    # ---- resident-input timing (eager) with live per-op / per-kernel events + clocks -------------
    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, f = sample_clocks_start(clk_path) if rank == 0 else (None, None)
    ops.PROFILE = {}
    lib.kon_profile_reset()
    lib.kon_profile_enable(1)
    l0 = lib.kon_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.kon_launch_count() - l0
    lib.kon_profile_enable(0)
    prof = ops.profile_summary()
    ops.PROFILE = None
    kprof = {kn: _lib.profile_read(kn) for kn in KERNEL_WORK}
    ms_eager = ms
    # ---- the short HBM-bound kernels again, back to back: a CUDA-event pair around ONE 50 us launch also
    # times ~10 us of launch latency, so the embedding kernels are timed as 4 x n_batches launches
    # between one pair of events, rotating over the distinct batches (436 MB of rows > L2) ---------
    iso = {}
    emb = model.sparse_embed
    if hasattr(emb, "plan"):
        emb = None                      # sharded: ids are exchanged first; skip the isolated pass
    if emb is not None:
        reps = 4 * len(resident)
        W = (26 * k + N_DENSE + 3) // 4 * 4
        xc = torch.empty((B, W), device=dev)
        outv = xc[:, :26 * k].view(B, 26, k)
        gv = torch.randn((B, W), device=dev)[:, :26 * k].view(B, 26, k)
        lib.kon_profile_reset()
        for kn, fn in (("embed_fwd_vec_kernel", lambda i: ops.embed_fwd_raw(emb.arena.detach(), resident[i % len(resident)][1],
                                                                            emb.field_row_offset, out=outv)),
                       ("embed_bwd", lambda i: ops.embed_bwd_raw(gv, resident[i % len(resident)][1],
                                                                 emb.field_row_offset, share_sort=False)),
                       # what the step's backward waits for: the routing was sorted at lookup time on the side stream
                       # (kon_embed_sort), the backward runs the segmented reduce + fixup only -- here with the
                       # first-order tables' gradient riding along (kon_embed_bwd_pair), as in the training step
                       ("embed_bwd_presorted", lambda i: ops.embed_bwd_raw(gv, resident[i % len(resident)][1],
                                                                           emb.field_row_offset, lin=g1v))):
            if kn == "embed_bwd_presorted":
                g1v = torch.randn((B, 1), device=dev).unsqueeze(1).expand(B, 26, 1)
                ops.new_step()
                for _, ids_r, _ in resident:
                    ops.embed_presort(ids_r, emb.field_row_offset)
                ops._join_side_streams()
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            # the launches are replayed from a CUDA graph: the Python/ctypes call (~50 us) is longer
            # than the kernel, so eager back-to-back launches would time the host, not the GPU
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for i in range(reps):
                        fn(i)
                gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                iso[kn] = e0.elapsed_time(e1) / reps
                del gr
            except Exception as e:      # noqa: BLE001
                print("isolated timing of", kn, "failed:", repr(e), file=sys.stderr)
                torch.cuda.synchronize()
            if kn == "embed_bwd_presorted":
                ops.end_step()
        del xc, outv, gv
    # ---- the same step replayed from a CUDA graph (kernel stats above come from the eager pass:
    # events cannot be recorded inside a capture).  `value` and `e2e` are each the MEDIAN of
    # args.windows windows of exactly args.steps steps: one 0.2 s window carries a few % of noise
    # (power-cap recovery), which is how a single-window e2e could come out faster than `value` --------
    graphed = False
    step_fn = step_resident
    if args.graph and (world == 1 or args.graph_multi):      # --no-graph-multi: eager steps when world_size > 1
        graphed = trainer.capture(*resident[0])
        if graphed:
            step_fn = step_resident_graph
            for i in range(3):
                step_fn(i)
        elif rank == 0:
            print("CUDA-graph capture failed, staying eager:", trainer.capture_error, file=sys.stderr)
    job.spin_up(step_fn, args.spinup)
    nvl0 = nvlink_counters(local) if (world > 1 and rank == 0) else None
    win_value = [timed(step_fn, args.steps) for _ in range(args.windows)]
    nvlink = None
    if nvl0 is not None:
        nvl1 = nvlink_counters(local)
        if nvl1 is not None:
            n_st = args.steps * args.windows
            k_emb = model.sparse_embed.dim if hasattr(model.sparse_embed, "dim") else 16
            frac = (world - 1) / world
            nvlink = {"source": "nvidia-smi nvlink -gt d (NVML data counters of rank 0's GPU, all links) around the timed windows",
                      "tx_bytes_per_step": (nvl1[0] - nvl0[0]) / n_st, "rx_bytes_per_step": (nvl1[1] - nvl0[1]) / n_st,
                      # per GPU and direction: embedding rows out + output-gradient rows back for the (N-1)/N of the
                      # global batch that lives on other ranks, the 4-byte ids, first-order terms; dense all-reduce apart
                      "algorithmic_exchange_bytes_per_step_per_direction":
                          frac * B * 26 * (2 * k_emb * 4 + 4 + 8)}
    ms = median(win_value)
    # ---- end-to-end timing (H2D + step + D2H) ---------------------------------------------
    for i in range(2):
        step_e2e(i)
    win_e2e = [timed(step_e2e, args.steps) for _ in range(args.windows)]
    ms_e2e = median(win_e2e)
    clocks = sample_clocks_stop(proc, f, clk_path, local) if rank == 0 else None
    exchange = "single GPU"
    if world > 1:
        exchange = ("fused: gather/scatter kernels store/load over NVLink peer memory (kon_embed_*_peer)"
                    if getattr(model.sparse_embed, "use_peer", False) else "NCCL all_to_all")
        job.check_peer()
    parallelism = "single GPU" if world == 1 else dctx.describe()
    job.close()
    del model, trainer, resident, stage

    # ---- the other BASELINE configs through the same trainer (compact) ------------------------------
    others = {}
    if args.other_models and name == "xdeepfm" and not args.big_tables and not strong:
        todo = [("deepfm", {}), ("dcn", {}), ("autoint", {})] if world == 1 else [("dcn", {})]
        for on, _ in todo:
            try:
                others[on] = quick_measure(on, args, dev, world, rank, args.batch, list(CRITEO_ROWS), mlp_dtype)
            except SystemExit:
                raise
            except Exception as e:      # noqa: BLE001
                others[on] = {"error": repr(e)}
                print("other_models:", on, "failed:", repr(e), file=sys.stderr)
                torch.cuda.synchronize()
        if world > 1:
            try:        # strong scaling: the global batch stays 65,536 (8192 per GPU at N = 8)
                o = quick_measure("xdeepfm", args, dev, world, rank, args.batch // world, list(CRITEO_ROWS), mlp_dtype)
                o["scaling"] = "strong"
                o["global_batch"] = args.batch // world * world
                others["xdeepfm_strong"] = o
            except Exception as e:      # noqa: BLE001
                others["xdeepfm_strong"] = {"error": repr(e)}
                torch.cuda.synchronize()
        if world == 8:
            try:        # BASELINE config 5: AutoInt + 8 x 100 M-row tables, row-wise sharded (51 GB of rows + Adam state)
                a5 = argparse.Namespace(**vars(args))
                a5.big_tables = "8x100000000"
                o = quick_measure("autoint", a5, dev, world, rank, args.batch, table_rows(a5), mlp_dtype)
                o["tables"] = "Criteo cardinalities with the 8 largest tables replaced by 100M-row tables (row-wise shards)"
                others["autoint_bigtables"] = o
            except Exception as e:      # noqa: BLE001
                others["autoint_bigtables"] = {"error": repr(e)}
                torch.cuda.synchronize()
    if rank != 0:
        return
    pk = peaks()
    work = algo_work(name, B, k)
    op_ms = {op: {"ms": mean_ms, "calls_per_step": calls / args.steps} for op, (calls, mean_ms) in prof.items()}
    kernels = {}
    for op, (calls, mean_ms) in prof.items():
        if op not in work or work[op][0] != "hbm":
            continue
        bound, amount = work[op]
        per_step_calls = calls / args.steps
        # ops called more than once per step (e.g. the two embedding arenas, stacked attention
        # layers) share the per-step algorithmic work evenly in this accounting
        rate = amount / (mean_ms * 1e-3 * per_step_calls) if op.startswith(("cross", "fm")) or per_step_calls <= 1 \
            else amount / (mean_ms * 1e-3)
        kernels[op] = {"bound": "hbm", "ms": mean_ms, "calls_per_step": per_step_calls,
                       "achieved": rate / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": rate / 1e9 / pk["hbm"]}
    # ---- per-kernel roofline: each main kernel is bracketed by CUDA events inside the library
    # (kon_profile_*), on the launching stream, inside the timed region ------------------------
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))
    kstats = {}
    for kn, (tot_ms, n) in kprof.items():
        if n == 0:
            continue
        op, share, per_launch = KERNEL_WORK[kn]
        if op not in work:
            continue
        bound, amount = work[op]
        lps = n / args.steps
        amount = amount * share * (lps if per_launch else 1.0)   # algorithmic work of this kernel family per step
        per_step_ms = tot_ms / args.steps
        rate = amount / (per_step_ms * 1e-3)
        peak = pk["hbm"] if bound == "hbm" else pk[TC_PEAK]
        scale = 1e9 if bound == "hbm" else 1e12
        kstats[kn] = {"bound": bound, "launches_per_step": lps, "ms_per_launch": per_step_ms / lps,
                      "work_per_launch": amount / lps, "achieved": rate / scale, "peak": peak,
                      "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": rate / scale / peak,
                      "share_of_step": per_step_ms / (ms_eager / args.steps)}
        if bound == "tensor":
            kstats[kn]["frac_of_sustained"] = rate / scale / pk["tc_sust"]
    for kn, ms_iso in iso.items():
        op = "embed_fwd" if kn == "embed_fwd_vec_kernel" else "embed_bwd"
        amount = work[op][1]
        kstats[kn + " (graph replay, back-to-back)"] = {
            "bound": "hbm", "launches_per_step": 1.0, "ms_per_launch": ms_iso, "work_per_launch": amount,
            "achieved": amount / (ms_iso * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
            "frac": amount / (ms_iso * 1e-3) / 1e9 / pk["hbm"], "share_of_step": ms_iso / (ms_eager / args.steps),
            "note": {"embed_bwd": "embed_bwd = routing (per-field counting sort: hist / scan / scatter per pass slot, run heads) + segmented sum + fixup; "
                                  "algorithmic bytes are the all-rows-unique worst case",
                     "embed_bwd_presorted": "the backward's critical path in the training step: segmented reduce + fixup of the "
                                            "embedding AND the first-order gradient in one pass (kon_embed_bwd_pair); the routing "
                                            "was sorted at lookup time on a side stream (kon_embed_sort)"}.get(
                         kn, "single launch, distinct id batches")}
    dom = max((kn for kn in kstats if "back-to-back" not in kn), key=lambda kn: kstats[kn]["share_of_step"], default=None)
    roof = None
    if dom is not None:
        kd = kstats[dom]
        tr = traffic_tab.get(dom, {}).get(name)
        roof = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"],
                "frac": kd["frac"], "traffic": tr,
                "peak_source": pk["src"] + (" (BURST bf16 peak: the kernel is timed in a sub-second window at boost "
                                            "clocks; frac_of_sustained beside it)" if kd["bound"] == "tensor" else ""),
                "launches_per_step": kd["launches_per_step"], "ms_per_launch": kd["ms_per_launch"],
                "work_per_launch": kd["work_per_launch"], "share_of_step": kd["share_of_step"],
                "note": "achieved = algorithmic work of the kernel's launches in a step / their summed CUDA-event time"}
        if "frac_of_sustained" in kd:
            roof["frac_of_sustained"] = kd["frac_of_sustained"]
    total = B * world * args.steps
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    # CPU baseline on rank 0, bounded sample
    cpu = None
    if not args.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N=1 only
        sb = CPU_SAMPLE_B[name]
        v, s_step, cores, note = cpu_arm(name, sb, args.cpu_steps, 1)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} steps of batch {sb} (oracle port of the reference layers -- pinned bit for bit to "
                         f"the reference's own source, tests/test_ref_pinned_cpu.py -- torch-CPU fp32, fwd+bwd+{note})"}
    line = {
        "metric": "train samples/s", "value": total / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16" if (name == "xdeepfm" and args.cin_precision == "bf16") else ("bf16" if mlp_dtype else "f32"),
        "data": "synthetic",
        "config": shared_config(name, args, world),
        "run": {"optimizer": "Adam (dense weights, fused) + row-wise LAZY Adam with lazy L2 on the touched embedding rows "
                             "(Keras applies both densely to every row; DESIGN.md section 7)",
                "cin": args.cin_precision + (" tcgen05, fp32 accumulate" if args.cin_precision == "bf16" else ""),
                "mlp_dtype": args.mlp_dtype,
                "cache": f"working set per step (>1 GB) exceeds the 126 MB L2; {args.n_batches} distinct batches rotate",
                "timing": f"median of {args.windows} windows of {args.steps} steps each (value and e2e alike), after --warmup "
                          f"steps and {args.spinup} s of untimed steps (clock ramp of ranks that idled during capture / verify)",
                "parallelism": parallelism, "embedding_exchange": exchange},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "h2d": ("double-buffered on a copy stream: the batch of step i+1 is copied while step i runs "
                        "(the reference's dataset.prefetch(2), DP:335-337)" if args.e2e_prefetch else "on the compute stream"),
                "ms_per_step": ms_e2e / args.steps, "windows_ms_per_step": [w / args.steps for w in win_e2e],
                "note": ("with the batch copy overlapped the end-to-end step costs what the resident step costs: the two "
                         "medians differ by window noise (+-0.5 %), in either direction")},
        "windows_ms_per_step": [w / args.steps for w in win_value],
        "gpu_launches": int(launches),
        "cuda_graph": graphed, "ms_per_step_eager": ms_eager / args.steps,
        "clocks": clocks,
        "nvlink": nvlink,
        "roofline": roof,
        "kernel_stats": kstats,
        "op_stats": kernels,
        "op_ms": op_ms,
        "other_models": others or None,
        "verified": verified,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def shutdown_process_group():
    """After the result line is out: tear the process group down, but never let a stuck NCCL teardown
    turn a finished measurement into a hung job."""
    try:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        import threading
        t = threading.Timer(45.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()
    except Exception as e:      # noqa: BLE001
        print("process-group teardown:", repr(e), file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="xdeepfm", choices=list(MODEL_CFG))
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cin-precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mlp-dtype", default="bf16", choices=["bf16", "f32", "tf32"])
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--no-e2e-prefetch", dest="e2e_prefetch", action="store_false",
                    help="end-to-end steps copy their batch on the compute stream (no overlap with the previous step)")
    ap.add_argument("--n-batches", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--windows", type=int, default=3, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--spinup", type=float, default=1.5,
                    help="seconds of untimed steps right before the timed windows (clock ramp; see Job.spin_up)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --batch is the GLOBAL batch, split over the ranks")
    ap.add_argument("--big-tables", default="", help="TxR: replace the T largest tables by R-row tables (config 5: 8x100000000)")
    ap.add_argument("--no-verify", dest="verify", action="store_false",
                    help="N > 1: skip the sharded-vs-single-GPU check that fills `verified`")
    ap.add_argument("--no-other-models", dest="other_models", action="store_false",
                    help="skip the compact DeepFM / DCN / AutoInt measurements in `other_models`")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="do not replay the step from a CUDA graph")
    ap.add_argument("--no-graph-multi", dest="graph_multi", action="store_false",
                    help="world_size > 1: do not capture the step (NCCL + peer kernels) in a CUDA graph")
    ap.add_argument("--graph-multi", dest="graph_multi", action="store_true", help=argparse.SUPPRESS)
    ap.set_defaults(graph_multi=True)
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: while the benchmark runs, file descriptor 1 points at
    # stderr (NCCL / C libraries print banners straight to fd 1), and is restored for the result.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    import io
    buf = io.StringIO()
    real_stdout, sys.stdout = sys.stdout, buf
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)
    finally:
        sys.stdout = real_stdout
        os.dup2(saved, 1)
        os.close(saved)
    out = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    for l in buf.getvalue().splitlines():
        if not l.startswith("{"):
            print(l, file=sys.stderr)
    if out:
        print(out[-1], flush=True)
    shutdown_process_group()


if __name__ == "__main__":
    main()
