mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r26_pytest.log
cat gpurun_out/r26_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r26_xdeepfm.json 2> gpurun_out/r26_bench.err
timeout 600 python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r26_deepfm.json 2>> gpurun_out/r26_bench.err
tail -3 gpurun_out/r26_bench.err
python - <<'PY'
import json
for m in ("xdeepfm","deepfm"):
    try:
        d=json.loads(open(f"gpurun_out/r26_{m}.json").read().strip().splitlines()[-1])
        print(m, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), json.dumps(d.get("roofline")), json.dumps(d.get("op_stats",{}).get("embed_bwd")), json.dumps(d.get("isolated")))
    except Exception as e: print(m, "ERR", e)
PY
