# round 2, call h: reuse-force test after the pair fix, PCIe duplex experiment, default bench line, train launch list
mkdir -p gpurun_out/r2h
timeout 600 python -m pytest tests/test_gpu_reuse_force.py -m gpu -q 2>&1 | tail -3
python profiles/exp_pcie_bidir.py | tee gpurun_out/r2h/pcie.json
timeout 900 python bench.py > gpurun_out/r2h/bench_default.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2h/bench_default.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2h/l2hmc_train_launches.csv python profiles/prof_l2hmc.py train 8 32 4 256 1 > gpurun_out/r2h/prof_train.log 2>&1; echo "train list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2h/l2hmc_train_launches.csv > gpurun_out/r2h/train_summary.md; head -45 gpurun_out/r2h/train_summary.md
