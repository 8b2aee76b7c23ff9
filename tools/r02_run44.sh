mkdir -p gpurun_out
for ov in 1 0 1 0; do
KON_OVERLAP_DENSE_OPT=$ov timeout 600 python bench.py --no-cpu-baseline --no-other-models 2>> gpurun_out/r44_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ov$ov', round(d['value']), d['ms_per_step'], d['windows_ms_per_step'])"
done
