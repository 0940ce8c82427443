# round 2, call ah: racecheck over the SU(3) / dense / training tests (the kernel-census tests use torch.profiler = CUPTI, which cannot attach next to the sanitizer: deselected)
mkdir -p gpurun_out/r2ah
T0=$(date +%s)
timeout 240 compute-sanitizer --tool racecheck --kernel-name kns=3l2b --print-limit 30 --log-file gpurun_out/r2ah/racecheck.log \
  python -m pytest tests/test_gpu_su3.py tests/test_gpu_dense.py tests/test_gpu_training.py tests/test_gpu_dynamics.py -q -m gpu -p no:cacheprovider -k "not library and not census" > gpurun_out/r2ah/pytest_under_racecheck.log 2>&1
echo "racecheck rc=$? $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r2ah/pytest_under_racecheck.log; tail -3 gpurun_out/r2ah/racecheck.log; grep -m 8 -B1 -A6 "hazard" gpurun_out/r2ah/racecheck.log | cut -c1-260 | head -60
# + the default bench line at HEAD (per-step timings of the secondary workloads)
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2ah/bench_default.log 2>&1; echo "bench rc=$? $(( $(date +%s) - T0 )) s"; grep '^{' gpurun_out/r2ah/bench_default.log > gpurun_out/r2ah/bench_default.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2ah/bench_default.jsonl').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
for k, v in d['secondary'].items(): print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'], v.get('ms_each_step'))
PY
grep -v '^{' gpurun_out/r2ah/bench_default.log | tail -4 | cut -c1-300
