mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bench_config_gpu.py tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -3
for m in deepfm dcn; do
for ov in 1 0; do
KON_OVERLAP_DENSE_OPT=$ov timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models > gpurun_out/r43_${m}_ov$ov.json 2>> gpurun_out/r43_bench.err
done; done
KON_OVERLAP_DENSE_OPT=1 timeout 600 python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r43_xdeepfm_ov1.json 2>> gpurun_out/r43_bench.err
tail -3 gpurun_out/r43_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r43_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"])
PY
