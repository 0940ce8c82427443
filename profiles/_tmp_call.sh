mkdir -p gpurun_out/r2n
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n/bench_n2.log 2>&1; echo rc=$?
grep '^{' gpurun_out/r2n/bench_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_GBps_per_gpu_all_ranks_copying'))
for k, v in d.get('secondary', {}).items(): print(k, v['ms_per_step'], v['value'], v.get('grad_allreduce'))
"
tail -5 gpurun_out/r2n/bench_n2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_train_check.py > gpurun_out/r2n/mgpu_check.log 2>&1; echo mgpu rc=$?; tail -6 gpurun_out/r2n/mgpu_check.log
