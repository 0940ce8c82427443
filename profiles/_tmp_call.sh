timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
