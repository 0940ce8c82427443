# round 2, call s: the driver's 8-GPU launch again, with the training step replayed from the split CUDA graphs
mkdir -p gpurun_out/r2s
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --thermalise-n 10 > gpurun_out/r2s/bench_n8.log 2>&1; echo rc=$?
grep '^{' gpurun_out/r2s/bench_n8.log > gpurun_out/r2s/bench_n8.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2s/bench_n8.jsonl').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k, v in d.get('secondary', {}).items(): print(k, v['ms_per_step'], v['value'], v.get('grad_allreduce'), v.get('cuda_graphs'))
PY
tail -2 gpurun_out/r2s/bench_n8.log | cut -c1-200
