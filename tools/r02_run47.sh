timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py tests/test_bench_config_gpu.py -m gpu -q -x -k "attn or attention or autoint or mha" 2>&1 | tail -3
timeout 600 python bench.py --model autoint --no-cpu-baseline --no-other-models 2>> gpurun_out/r46_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('autoint', round(d['value']), d['ms_per_step'], d['windows_ms_per_step'], d['op_ms']['attn_bwd'], d['op_ms']['attn_fwd'])"
