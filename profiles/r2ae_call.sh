# round 2, call ae: the driver's 2-GPU launch of the default bench at HEAD + the 2-GPU gradient check
mkdir -p gpurun_out/r2ae
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ae/bench_n2.log 2>&1; echo "bench rc=$? $(( $(date +%s) - T0 )) s"; grep '^{' gpurun_out/r2ae/bench_n2.log > gpurun_out/r2ae/bench_n2.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ae/bench_n2.jsonl').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], 'launches', d['gpu_launches'])
for k, v in d['secondary'].items(): print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'], v.get('cuda_graphs'), v['gpu_launches'], v.get('grad_allreduce'))
PY
grep -v '^{' gpurun_out/r2ae/bench_n2.log | tail -5 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tests/mgpu_train_check.py 2>&1 | tail -4 | cut -c1-300
