timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -v Warning | tail -40
