timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "census or library" 2>&1 | tail -15
