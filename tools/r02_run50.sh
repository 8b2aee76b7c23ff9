timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py tests/test_bench_config_gpu.py -m gpu -q -x -k "attn or attention or autoint or mha" 2>&1 | grep -E "^E |FAILED|passed|failed" | head -12
echo "---- PF off"
KON_ATTN_PF_OFF=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_ref_pinned_gpu.py tests/test_bench_config_gpu.py -m gpu -q -x -k "attn or attention or autoint or mha" 2>&1 | grep -E "^E |FAILED|passed|failed" | head -12
