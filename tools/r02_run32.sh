timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -5
