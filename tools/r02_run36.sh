mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for m in deepfm dcn; do
for acc in 1 0; do
KON_ACC_XGRAD=$acc timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models > gpurun_out/r36_${m}_acc$acc.json 2>> gpurun_out/r36_bench.err
done; done
tail -3 gpurun_out/r36_bench.err
python - <<'PY'
import json
for m in ("deepfm","dcn"):
  for acc in (1,0):
    d=json.loads(open(f"gpurun_out/r36_{m}_acc{acc}.json").read().strip().splitlines()[-1])
    print(m, "acc",acc, round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"], round(d["e2e"]["value"]))
PY
