timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -q -x -k "accumulated_in_place" 2>&1 | tail -5
