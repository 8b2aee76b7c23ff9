mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bench_config_gpu.py tests/test_models_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -3
for m in deepfm dcn autoint; do
for ov in 1 0; do
KON_OVERLAP_DENSE_OPT=$ov timeout 600 python bench.py --model $m --no-cpu-baseline --no-other-models 2>> gpurun_out/r45_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m ov$ov', round(d['value']), d['ms_per_step'], d['windows_ms_per_step'])"
done; done
