# round 2, call p: suite at HEAD, default bench line (both arms), GEMM shapes, launch lists, ncu of the general GEMM and the step kernel
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2p/pytest_all.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/r2p/bench_reference.jsonl; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/r2p/bench_default.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2p/bench_default.log > gpurun_out/r2p/bench_default.jsonl; cut -c1-300 gpurun_out/r2p/bench_default.jsonl
python profiles/bench_gemm.py > gpurun_out/r2p/bench_gemm.jsonl 2>&1; cat gpurun_out/r2p/bench_gemm.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p/launch_list.csv python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 1 > gpurun_out/r2p/bench_under_ncu.log 2>&1; echo "list rc=$?"
python profiles/summarize_launches.py gpurun_out/r2p/launch_list.csv > gpurun_out/r2p/launch_list_summary.md; head -16 gpurun_out/r2p/launch_list_summary.md
timeout 600 ncu --set full --clock-control none -k regex:k_gemm_bf16 -s 8 -c 1 -f -o gpurun_out/r2p/gemm_dz python profiles/bench_gemm.py dz > /dev/null 2>&1; python profiles/summarize_ncu.py gpurun_out/r2p/gemm_dz.ncu-rep > gpurun_out/r2p/gemm_dz_ncu_full.md; grep "|" gpurun_out/r2p/gemm_dz_ncu_full.md | head -24
timeout 600 ncu --set full --clock-control none -k regex:k_gemm_bf16 -s 8 -c 1 -f -o gpurun_out/r2p/gemm_dwh python profiles/bench_gemm.py dw_heads > /dev/null 2>&1; python profiles/summarize_ncu.py gpurun_out/r2p/gemm_dwh.ncu-rep > gpurun_out/r2p/gemm_dw_heads_ncu_full.md; grep "|" gpurun_out/r2p/gemm_dw_heads_ncu_full.md | head -8
rm -f gpurun_out/r2p/*.ncu-rep
