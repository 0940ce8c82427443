# round 2, call b: suite at HEAD (no opt-ins left), the full default bench line, ncu launch list + full capture of the step kernel
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q -rs --durations=5 > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b/pytest.log
( time timeout 900 python bench.py > gpurun_out/r2b/bench_default.log 2> gpurun_out/r2b/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2b/bench_default.log; tail -5 gpurun_out/r2b/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2b/bench_reference.log 2>&1; echo "ref arm rc=$?"; cut -c1-300 gpurun_out/r2b/bench_reference.log | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b/launches.csv python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 3 > gpurun_out/r2b/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_ep -s 12 -c 1 -f -o gpurun_out/r2b/force_ep_r2 python profiles/prof_su3.py 16 64 10 2 > gpurun_out/r2b/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_epx -c 3 -f -o gpurun_out/r2b/force_epx_r2 python profiles/prof_su3.py 16 64 10 1 > gpurun_out/r2b/ncu_full_x.log 2>&1; echo "ncu full x rc=$?"
for w in su3_8x8x8x8_nb256_l2hmc_eval_bf16; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --cuda-graphs > gpurun_out/r2b/bench_${w}_graphs.log 2>&1; echo "$w graphs rc=$?"; cut -c1-200 gpurun_out/r2b/bench_${w}_graphs.log | tail -1
done
