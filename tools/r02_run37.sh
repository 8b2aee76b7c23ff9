mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "cross or dcn or fm" 2>&1 | tail -3
for acc in 1 0; do
KON_ACC_XGRAD=$acc timeout 600 python bench.py --model dcn --no-cpu-baseline --no-other-models > gpurun_out/r37_dcn_acc$acc.json 2>> gpurun_out/r37_bench.err
done
python - <<'PY'
import json
for acc in (1,0):
    d=json.loads(open(f"gpurun_out/r37_dcn_acc{acc}.json").read().strip().splitlines()[-1])
    print("dcn acc",acc, round(d["value"]), d["ms_per_step"], d["op_ms"]["cross_bwd"]["ms"])
PY
