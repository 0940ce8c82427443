# round 2, call d: the default bench line (both arms), launch list, ONE full capture of the step kernel summarised on the box
# (the .ncu-rep files are ~38 MB each: gpurun_out is capped at 64 MiB, so only the summaries travel back)
mkdir -p gpurun_out/r2d
timeout 900 python bench.py > gpurun_out/r2d/bench_default.log 2> gpurun_out/r2d/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2d/bench_reference.log 2>&1; echo "ref arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d/launches.csv python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2d/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_ep -s 12 -c 1 -f -o /tmp/force_ep_r2 python profiles/prof_su3.py 16 64 10 2 > gpurun_out/r2d/ncu_full.log 2>&1; echo "ncu full rc=$?"
python profiles/summarize_ncu.py /tmp/force_ep_r2.ncu-rep > gpurun_out/r2d/force_ep_ncu_full.md 2>&1
ncu -i /tmp/force_ep_r2.ncu-rep --page source --csv > gpurun_out/r2d/force_ep_source.csv 2>/dev/null; ls -la gpurun_out/r2d/force_ep_source.csv
ncu -i /tmp/force_ep_r2.ncu-rep --page details --csv > gpurun_out/r2d/force_ep_details.csv 2>/dev/null
L2B_SU3_FORCE_VARIANT=32 timeout 600 ncu --set full --clock-control none -k regex:k_force_tma -s 6 -c 1 -f -o /tmp/force_tma_r2 python profiles/prof_su3.py 16 64 10 2 > gpurun_out/r2d/ncu_full_tma.log 2>&1; echo "ncu tma rc=$?"
python profiles/summarize_ncu.py /tmp/force_tma_r2.ncu-rep > gpurun_out/r2d/force_tma_ncu_full.md 2>&1
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline --cuda-graphs > gpurun_out/r2d/bench_eval_graphs.log 2>&1; echo "eval graphs rc=$?"
timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline > gpurun_out/r2d/bench_eval.log 2>&1; echo "eval rc=$?"
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline > gpurun_out/r2d/bench_train.log 2>&1; echo "train rc=$?"
timeout 300 python -m pytest tests/test_gpu_dynamics.py -m gpu -q -x -k "beyond_one_block" > gpurun_out/r2d/pytest_u1big.log 2>&1; echo "pytest u1 big rc=$?"; tail -3 gpurun_out/r2d/pytest_u1big.log
du -sh gpurun_out/r2d
