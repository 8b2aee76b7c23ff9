mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py tests/test_bench_config_gpu.py -m gpu -q -x -k "attn or attention or autoint or mha" 2>&1 | tail -2
for i in 1 2; do
for off in "" 1; do
KON_ATTN_PF_OFF=$off; [ -z "$off" ] && unset KON_ATTN_PF_OFF || export KON_ATTN_PF_OFF
timeout 600 python bench.py --model autoint --no-cpu-baseline --no-other-models 2>> gpurun_out/ab_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); om=d['op_ms']; print('autoint pf_off=[$off]', round(d['value']), d['ms_per_step'], {k:round(v['ms'],4) for k,v in om.items() if 'attn' in k})"
done; done
