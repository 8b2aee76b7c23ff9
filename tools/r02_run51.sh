mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cin_tc_gpu.py tests/test_bench_config_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "cin or xdeepfm" 2>&1 | tail -2
for i in 1 2; do
for v in b200 vB; do
KON_B200_LIB=$PWD/ml_function_b200/libkon_$v.so timeout 600 python bench.py --no-cpu-baseline --no-other-models 2>> gpurun_out/ab_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('xdeepfm $v', round(d['value']), d['ms_per_step'], d['op_ms']['cin_bwd']['ms'], d['op_ms']['cin_fwd']['ms'])"
done; done
