# round 2, call ac: smoke, reference arm, default bench line at HEAD (cfg 1 in `secondary` and in the reference arm), with wall times
mkdir -p gpurun_out/r2ac
T0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; echo "smoke $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
timeout 900 python bench.py --impl reference 2>gpurun_out/r2ac/ref.err | grep '^{' > gpurun_out/r2ac/bench_reference.jsonl; echo "ref rc=$? $(( $(date +%s) - T0 )) s"; cut -c1-400 gpurun_out/r2ac/bench_reference.jsonl; tail -2 gpurun_out/r2ac/ref.err
T0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2ac/bench_default.log 2>&1; echo "bench rc=$? $(( $(date +%s) - T0 )) s"; grep '^{' gpurun_out/r2ac/bench_default.log > gpurun_out/r2ac/bench_default.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ac/bench_default.jsonl').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
print('thermalised', d['thermalised']['ms_per_step'], 'gpu_reference', d['gpu_reference']['value'], 'cpu', d['cpu_baseline']['value'])
for k, v in d['secondary'].items(): print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'], v.get('cuda_graphs'), v['gpu_launches'])
PY
grep -v '^{' gpurun_out/r2ac/bench_default.log | tail -5 | cut -c1-300
