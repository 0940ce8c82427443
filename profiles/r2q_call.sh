# round 2, call q: the driver's 8-GPU launch of bench.py (headline + secondary incl. the cfg-5 training step with the bucketed all-reduce)
mkdir -p gpurun_out/r2q
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --thermalise-n 10 > gpurun_out/r2q/bench_n8.log 2>&1; echo rc=$?
grep '^{' gpurun_out/r2q/bench_n8.log > gpurun_out/r2q/bench_n8.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2q/bench_n8.jsonl').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_GBps_per_gpu_all_ranks_copying'), d['config'].get('host_cores_bound'))
for k, v in d.get('secondary', {}).items(): print(k, v['ms_per_step'], v['value'], v.get('grad_allreduce'))
PY
tail -3 gpurun_out/r2q/bench_n8.log | cut -c1-300
