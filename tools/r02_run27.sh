mkdir -p gpurun_out
for v in b200 v3; do
KON_B200_LIB=$PWD/ml_function_b200/libkon_$v.so timeout 600 python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r27_deepfm_$v.json 2>> gpurun_out/r27_bench.err
done
python - <<'PY'
import json
for v in ("b200","v3"):
    d=json.loads(open(f"gpurun_out/r27_deepfm_{v}.json").read().strip().splitlines()[-1])
    print(v, d["value"], d["ms_per_step"], d["windows_ms_per_step"], d["kernel_stats"]["embed_reduce_kernel"]["ms_per_launch"], d["op_stats"]["embed_bwd"]["ms"])
PY
