# round 2, call al (2 GPUs): one-pass hand-back of the gradient bucket -- 2-GPU gradient check and the graphed training step
mkdir -p gpurun_out/r2al
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_train_check.py 2>&1 | grep -E "trainer path|graphed|Error|error" | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>gpurun_out/r2al/train.err | grep '^{' > gpurun_out/r2al/train_n2.jsonl; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2al/train_n2.jsonl; tail -2 gpurun_out/r2al/train.err | cut -c1-200
