# round 2, call aj (2 GPUs): NCCL picked RING_LL for the 362 MB gradient bucket (1.15 ms = 314 GB/s); same step with NCCL_PROTO=Simple
mkdir -p gpurun_out/r2aj
for proto in default Simple; do
  if [ $proto = default ]; then unset NCCL_PROTO; else export NCCL_PROTO=$proto; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 profiles/prof_l2hmc.py train 8 32 4 256 10 --graph --table > gpurun_out/r2aj/train_n2_$proto.txt 2>&1; echo "$proto rc=$?"
  grep "ms/call" gpurun_out/r2aj/train_n2_$proto.txt; grep -E "ncclDev" gpurun_out/r2aj/train_n2_$proto.txt | cut -c1-72,150-240
done
