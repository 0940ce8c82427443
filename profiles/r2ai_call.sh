# round 2, call ai (2 GPUs): where the multi-rank training step spends its extra 3.8 ms -- torch profiler table of one graphed step on rank 0
mkdir -p gpurun_out/r2ai
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 profiles/prof_l2hmc.py train 8 32 4 256 5 --graph --table > gpurun_out/r2ai/train_n2_table.txt 2>&1; echo "rc=$?"
grep "ms/call" gpurun_out/r2ai/train_n2_table.txt; grep -E "^ *(Name|nccl|ncclDev|void|l2b|Memcpy|Memset|aten)" gpurun_out/r2ai/train_n2_table.txt | cut -c1-72,150-240 | head -45
tail -3 gpurun_out/r2ai/train_n2_table.txt | cut -c1-200
