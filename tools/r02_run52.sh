mkdir -p gpurun_out
KON_B200_LIB=$PWD/ml_function_b200/libkon_vB.so timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_peer_gpu.py tests/test_ref_pinned_gpu.py -m gpu -q -x -k "embed or sparse or peer" 2>&1 | tail -2
for i in 1 2; do
for v in b200 vB; do
KON_B200_LIB=$PWD/ml_function_b200/libkon_$v.so timeout 600 python bench.py --model deepfm --no-cpu-baseline --no-other-models 2>> gpurun_out/ab_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['kernel_stats']; print('deepfm $v', round(d['value']), d['ms_per_step'], round(d['op_ms']['embed_fwd']['ms'],4), {k:round(v['ms_per_launch'],4) for k,v in ks.items() if 'embed_fwd' in k})"
done; done
