# round 2, call g (re-entry after the container was re-created): full GPU suite at HEAD, eval/train lines, launch list, input GEMM ncu
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/r2g/pytest_all.log
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline > gpurun_out/r2g/bench_eval.log 2>&1; echo "eval rc=$?"; grep '^{' gpurun_out/r2g/bench_eval.log | cut -c1-250
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline --cuda-graphs > gpurun_out/r2g/bench_eval_graphs.log 2>&1; echo "eval graphs rc=$?"; grep '^{' gpurun_out/r2g/bench_eval_graphs.log | cut -c1-250
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline > gpurun_out/r2g/bench_train.log 2>&1; echo "train rc=$?"; grep '^{' gpurun_out/r2g/bench_train.log | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2g/l2hmc_eval_launches.csv python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2g/prof_eval.log 2>&1; echo "eval list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2g/l2hmc_eval_launches.csv | head -30
timeout 600 ncu --set full --clock-control none -k regex:k_su3_input_gemm -s 2 -c 1 -f -o /tmp/input_gemm python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2g/ncu_input.log 2>&1; echo "ncu input rc=$?"
python profiles/summarize_ncu.py /tmp/input_gemm.ncu-rep > gpurun_out/r2g/input_gemm_ncu_full.md 2>&1; grep "|" gpurun_out/r2g/input_gemm_ncu_full.md | head -24
