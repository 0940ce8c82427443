timeout 600 python -m pytest tests/test_gpu_reuse_force.py -m gpu -q -k bit_identical 2>&1 | grep -v Warning | tail -40
