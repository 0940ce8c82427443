# round 2, call u (re-entry): full GPU suite at HEAD, smoke, default bench line, launch lists of the cfg-5 training / eval steps, ncu of the new kernels
mkdir -p gpurun_out/r2u
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2u/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2u/bench_default.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2u/bench_default.log > gpurun_out/r2u/bench_default.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2u/bench_default.jsonl').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'], 'parity', d['parity']['ok'], 'launches', d['gpu_launches'])
print('thermalised', d['thermalised']['ms_per_step'], 'gpu_reference', d['gpu_reference']['value'], 'cpu', d['cpu_baseline']['value'])
for k, v in d['secondary'].items(): print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'], v.get('cuda_graphs'), v['gpu_launches'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2u/l2hmc_train_launch_list.csv python profiles/prof_l2hmc.py train 8 32 4 256 1 > gpurun_out/r2u/prof_train.log 2>&1; echo "train list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2u/l2hmc_train_launch_list.csv > gpurun_out/r2u/l2hmc_train_launch_list_summary.md; head -40 gpurun_out/r2u/l2hmc_train_launch_list_summary.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2u/l2hmc_eval_launch_list.csv python profiles/prof_l2hmc.py eval 8 256 4 256 1 > gpurun_out/r2u/prof_eval.log 2>&1; echo "eval list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2u/l2hmc_eval_launch_list.csv > gpurun_out/r2u/l2hmc_eval_launch_list_summary.md; head -24 gpurun_out/r2u/l2hmc_eval_launch_list_summary.md
timeout 300 ncu --set full --clock-control none -k regex:k_heads_vupdate_bwd -s 4 -c 1 -f -o /tmp/hvb python profiles/prof_l2hmc.py train 8 32 4 256 1 > /dev/null 2>&1; python profiles/summarize_ncu.py /tmp/hvb.ncu-rep > gpurun_out/r2u/heads_vupdate_bwd_ncu_full.md 2>&1; grep "|" gpurun_out/r2u/heads_vupdate_bwd_ncu_full.md | head -24
timeout 300 ncu --set full --clock-control none -k regex:k_update_gauge_bwd -s 4 -c 1 -f -o /tmp/ugb python profiles/prof_l2hmc.py train 8 32 4 256 1 > /dev/null 2>&1; python profiles/summarize_ncu.py /tmp/ugb.ncu-rep > gpurun_out/r2u/update_gauge_bwd_ncu_full.md 2>&1; grep "|" gpurun_out/r2u/update_gauge_bwd_ncu_full.md | head -24
