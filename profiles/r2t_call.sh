# round 2, call t: the driver's sequence at HEAD -- GPU suite, smoke, reference arm, default bench
mkdir -p gpurun_out/r2t
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2t/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2t/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/r2t/bench_reference.jsonl; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/r2t/bench_default.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2t/bench_default.log > gpurun_out/r2t/bench_default.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2t/bench_default.jsonl').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
print('thermalised', d['thermalised']['ms_per_step'], 'gpu_reference', d['gpu_reference']['value'], 'cpu', d['cpu_baseline']['value'])
for k, v in d['secondary'].items(): print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'], v.get('cuda_graphs'))
PY
