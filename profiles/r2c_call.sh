# round 2, call c: default bench line (full), launch list + one full capture of the step kernel, force variants incl. the TMA-staged ones
mkdir -p gpurun_out/r2c
timeout 300 python -m pytest tests/test_gpu_su3.py -m gpu -q -x -k "tma or folded" > gpurun_out/r2c/pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -3 gpurun_out/r2c/pytest_tma.log
( time timeout 900 python bench.py > gpurun_out/r2c/bench_default.log 2> gpurun_out/r2c/bench_default.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/r2c/bench_default.log; tail -3 gpurun_out/r2c/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2c/bench_reference.log 2>&1; echo "ref arm rc=$?"
timeout 600 python profiles/tune_force.py 22 26 32 33 29 > gpurun_out/r2c/tune_force.jsonl 2>&1; echo "tune rc=$?"; cat gpurun_out/r2c/tune_force.jsonl | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2c/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_ep -s 12 -c 1 -f -o gpurun_out/r2c/force_ep_r2 python profiles/prof_su3.py 16 64 10 2 > gpurun_out/r2c/ncu_full.log 2>&1; echo "ncu full rc=$?"
L2B_SU3_FORCE_VARIANT=32 timeout 600 ncu --set full --clock-control none -k regex:k_force_tma -s 6 -c 1 -f -o gpurun_out/r2c/force_tma_r2 python profiles/prof_su3.py 16 64 10 2 > gpurun_out/r2c/ncu_full_tma.log 2>&1; echo "ncu tma rc=$?"
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline --cuda-graphs > gpurun_out/r2c/bench_eval_graphs.log 2>&1; echo "eval graphs rc=$?"
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline > gpurun_out/r2c/bench_eval.log 2>&1; echo "eval rc=$?"
ls -la gpurun_out/r2c
