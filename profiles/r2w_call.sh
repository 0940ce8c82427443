# round 2, call w: column-wise exp adjoint / two-phase k_update_gauge_bwd: training tests, register-cap A/B, train step, cfg-1 workloads
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_trainer.py tests/test_gpu_graphs.py -x -q -m gpu 2>&1 | tail -3
python profiles/time_gauge_bwd.py 2>&1 | tail -5
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>/dev/null | grep '^{' > gpurun_out/r2w/bench_train.jsonl; cut -c1-260 gpurun_out/r2w/bench_train.jsonl
for w in u1_16x16_nb128_l2hmc_eval_f32 u1_16x16_nb128_l2hmc_train_f32; do timeout 600 python bench.py --workload $w --no-cpu-baseline 2>gpurun_out/r2w/$w.err | grep '^{' > gpurun_out/r2w/$w.jsonl; cut -c1-330 gpurun_out/r2w/$w.jsonl; tail -3 gpurun_out/r2w/$w.err; done
