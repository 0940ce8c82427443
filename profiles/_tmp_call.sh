mkdir -p gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_trainer.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>&1 | grep -E '^\{|Error|error' | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2l/l2hmc_eval_launches.csv python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2l/prof_eval.log 2>&1; echo "eval list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2l/l2hmc_eval_launches.csv > gpurun_out/r2l/eval_summary.md; head -20 gpurun_out/r2l/eval_summary.md; tail -2 gpurun_out/r2l/eval_summary.md
