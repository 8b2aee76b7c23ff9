mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for m in deepfm dcn xdeepfm; do
timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models > gpurun_out/r42_${m}.json 2>> gpurun_out/r42_bench.err
done
tail -3 gpurun_out/r42_bench.err
python - <<'PY'
import json
for m in ("deepfm","dcn","xdeepfm"):
    d=json.loads(open(f"gpurun_out/r42_{m}.json").read().strip().splitlines()[-1])
    print(m, round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"], round(d["e2e"]["value"]), d["op_ms"]["embed_fwd"])
PY
