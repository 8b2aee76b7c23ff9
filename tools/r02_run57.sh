timeout 900 python -m pytest tests/test_models_gpu.py tests/test_bench_config_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --model deepfm --no-cpu-baseline --no-other-models 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deepfm', round(d['value']), d['ms_per_step'])"
