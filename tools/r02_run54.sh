mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py -m gpu -q -x -k "cross or dcn" 2>&1 | tail -2
bash tools/r02_ab.sh dcn
