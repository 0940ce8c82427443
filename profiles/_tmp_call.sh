timeout 900 python -m pytest tests/test_gpu_vnet.py tests/test_gpu_training.py tests/test_gpu_trainer.py tests/test_gpu_graphs.py tests/test_gpu_dense.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>/dev/null | grep '^{' | cut -c1-230
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline 2>/dev/null | grep '^{' | cut -c1-230
