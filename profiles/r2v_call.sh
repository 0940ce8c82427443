# round 2, call v: exp adjoint on the scalar Cayley-Hamilton recurrence (k_update_gauge_bwd): training tests, train step, ncu of the kernel
mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_trainer.py tests/test_gpu_graphs.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>/dev/null | grep '^{' > gpurun_out/r2v/bench_train.jsonl; cut -c1-260 gpurun_out/r2v/bench_train.jsonl
timeout 300 ncu --set full --clock-control none -k regex:k_update_gauge_bwd -s 4 -c 1 -f -o /tmp/ugb python profiles/prof_l2hmc.py train 8 32 4 256 1 > /dev/null 2>&1; python profiles/summarize_ncu.py /tmp/ugb.ncu-rep > gpurun_out/r2v/update_gauge_bwd_ncu_full.md 2>&1; grep "|" gpurun_out/r2v/update_gauge_bwd_ncu_full.md | head -24
