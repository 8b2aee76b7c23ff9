set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02o_pytest.log
tail -3 gpurun_out/r02o_pytest.log
for v in b200 v1 v2 v3; do
KON_B200_LIB=$PWD/ml_function_b200/libkon_$v.so python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r02o_deepfm_$v.json 2>> gpurun_out/r02o_bench.err
done
KON_FUSE_LIN=0 python bench.py --model deepfm --no-cpu-baseline --no-other-models > gpurun_out/r02o_deepfm_nofuse.json 2>> gpurun_out/r02o_bench.err
