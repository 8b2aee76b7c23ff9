# same-box A/B of two builds of the library: KON_B200_LIB picks the .so
mkdir -p gpurun_out
M=${1:-autoint}
for i in 1 2; do
for v in b200 vB; do
KON_B200_LIB=$PWD/ml_function_b200/libkon_$v.so timeout 600 python bench.py --model $M --no-cpu-baseline --no-other-models 2>> gpurun_out/ab_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); om=d['op_ms']; print('$M $v', round(d['value']), d['ms_per_step'], {k:round(v['ms'],4) for k,v in om.items() if 'attn' in k or 'cross' in k or 'embed' in k})"
done; done
