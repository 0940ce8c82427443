# round 2, call f: the tensor-core input layer + paired updates (tests, eval-step timing, launch list)
mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_vnet.py tests/test_gpu_dynamics.py tests/test_gpu_reuse_force.py tests/test_gpu_graphs.py -m gpu -q > gpurun_out/r2f/pytest_vnet.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2f/pytest_vnet.log
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline > gpurun_out/r2f/bench_eval.log 2>&1; echo "eval rc=$?"; grep '^{' gpurun_out/r2f/bench_eval.log | cut -c1-250
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline --cuda-graphs > gpurun_out/r2f/bench_eval_graphs.log 2>&1; echo "eval graphs rc=$?"; grep '^{' gpurun_out/r2f/bench_eval_graphs.log | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f/l2hmc_eval_launches.csv python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2f/prof_eval.log 2>&1; echo "eval list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2f/l2hmc_eval_launches.csv | head -24
timeout 600 ncu --set full --clock-control none -k regex:k_su3_input_gemm -s 2 -c 1 -f -o /tmp/input_gemm python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2f/ncu_input.log 2>&1; echo "ncu input rc=$?"
python profiles/summarize_ncu.py /tmp/input_gemm.ncu-rep > gpurun_out/r2f/input_gemm_ncu_full.md 2>&1; grep "|" gpurun_out/r2f/input_gemm_ncu_full.md | head -24
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2f/pytest_all.log
