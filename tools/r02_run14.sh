set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02n_pytest.log
python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r02n_deepfm.json 2>> gpurun_out/r02n_bench.err
KON_FUSE_LIN=0 python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r02n_deepfm_nofuse.json 2>> gpurun_out/r02n_bench.err
tail -20 gpurun_out/r02n_pytest.log
