set -x
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_bench_config_gpu.py tests/test_ref_pinned_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02r_pytest.log
tail -3 gpurun_out/r02r_pytest.log
python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r02r_x_presort.json 2>> gpurun_out/r02r_bench.err
KON_PRESORT_X=0 python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r02r_x_nopresort.json 2>> gpurun_out/r02r_bench.err
python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r02r_x_presort2.json 2>> gpurun_out/r02r_bench.err
KON_PRESORT_X=0 python bench.py --no-cpu-baseline --no-other-models > gpurun_out/r02r_x_nopresort2.json 2>> gpurun_out/r02r_bench.err
