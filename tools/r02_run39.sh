mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bench_config_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py -m gpu -q -x 2>&1 | tail -4
for m in deepfm dcn xdeepfm; do
timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models > gpurun_out/r39_${m}.json 2>> gpurun_out/r39_bench.err
done
tail -3 gpurun_out/r39_bench.err
python - <<'PY'
import json
for m in ("deepfm","dcn","xdeepfm"):
    d=json.loads(open(f"gpurun_out/r39_{m}.json").read().strip().splitlines()[-1])
    print(m, round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"])
PY
