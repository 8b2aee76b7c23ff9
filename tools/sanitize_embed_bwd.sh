#!/bin/bash
# compute-sanitizer over the embedding-backward tests (routing kernels: shared-memory counters, warp-synchronous
# read-modify-write; segmented sum: per-warp staging + CTA stitch).  One GPU.  Logs under gpurun_out/.
mkdir -p gpurun_out
T="tests/test_kernels_gpu.py -k embed_bwd -m gpu -q -x -p no:cacheprovider"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T > gpurun_out/sanitize_${tool}.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_${tool}.log | tail -3
done
