mkdir -p gpurun_out
for m in xdeepfm deepfm; do
timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models > gpurun_out/r33_$m.json 2>> gpurun_out/r33_bench.err
done
timeout 600 python bench.py --model deepfm --no-cpu-baseline --no-other-models --no-e2e-prefetch > gpurun_out/r33_deepfm_noprefetch.json 2>> gpurun_out/r33_bench.err
tail -3 gpurun_out/r33_bench.err
python - <<'PY'
import json
for m in ("xdeepfm","deepfm","deepfm_noprefetch"):
    d=json.loads(open(f"gpurun_out/r33_{m}.json").read().strip().splitlines()[-1])
    print(m, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["e2e"]["windows_ms_per_step"])
PY
