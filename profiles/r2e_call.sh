# round 2, call e (2 GPUs): multi-rank training check (bucketed gradient exchange, rank-0 broadcast), the full bench line at N=2,
# and launch lists of one L2HMC eval / train step on one GPU
mkdir -p gpurun_out/r2e
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_train_check.py > gpurun_out/r2e/mgpu_train_check.log 2>&1; echo "mgpu check rc=$?"; grep -v Warning gpurun_out/r2e/mgpu_train_check.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e/bench_n2.log 2> gpurun_out/r2e/bench_n2.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/r2e/bench_n2.log | cut -c1-250; tail -3 gpurun_out/r2e/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2e/l2hmc_eval_launches.csv python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2e/prof_eval.log 2>&1; echo "eval list rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2e/l2hmc_train_launches.csv python profiles/prof_l2hmc.py train 8 32 4 256 1 > gpurun_out/r2e/prof_train.log 2>&1; echo "train list rc=$?"
du -sh gpurun_out/r2e
