# round 2, call ak: initcheck (uninitialised device-memory reads) over the libl2b kernels; ncu --set full of the step kernel and of the final half kick at HEAD
mkdir -p gpurun_out/r2ak
T0=$(date +%s)
timeout 200 compute-sanitizer --tool initcheck --kernel-name kns=3l2b --print-limit 40 --log-file gpurun_out/r2ak/initcheck.log \
  python -m pytest tests/test_gpu_su3.py tests/test_gpu_u1.py tests/test_gpu_vnet.py tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_dense.py tests/test_gpu_training.py tests/test_gpu_dynamics.py -q -m gpu -p no:cacheprovider -k "not library and not census" > gpurun_out/r2ak/pytest_under_initcheck.log 2>&1
echo "initcheck rc=$? $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r2ak/pytest_under_initcheck.log; tail -3 gpurun_out/r2ak/initcheck.log; grep -m 6 -A8 "Uninitialized" gpurun_out/r2ak/initcheck.log | cut -c1-240 | head -70
T0=$(date +%s)
timeout 300 ncu --set full --clock-control none -k regex:k_force_ep\< -s 6 -c 1 -f -o /tmp/fep python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /dev/null 2>&1; python profiles/summarize_ncu.py /tmp/fep.ncu-rep > gpurun_out/r2ak/force_ep_ncu_full.md 2>&1; grep "|" gpurun_out/r2ak/force_ep_ncu_full.md | head -26
timeout 300 ncu --set full --clock-control none -k regex:k_force_epx -s 5 -c 1 -f -o /tmp/fepx python bench.py --headline-only --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /dev/null 2>&1; python profiles/summarize_ncu.py /tmp/fepx.ncu-rep > gpurun_out/r2ak/force_epx_ncu_full.md 2>&1; grep "|" gpurun_out/r2ak/force_epx_ncu_full.md | head -26
echo "ncu $(( $(date +%s) - T0 )) s"
