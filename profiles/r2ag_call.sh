# round 2, call ag: compute-sanitizer racecheck (shared-memory hazards) and synccheck over the libl2b kernels
mkdir -p gpurun_out/r2ag
T0=$(date +%s)
timeout 270 compute-sanitizer --tool racecheck --kernel-name kns=3l2b --print-limit 30 --log-file gpurun_out/r2ag/racecheck.log \
  python -m pytest tests/test_gpu_u1.py tests/test_gpu_vnet.py tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_su3.py tests/test_gpu_dense.py -q -m gpu -p no:cacheprovider > gpurun_out/r2ag/pytest_under_racecheck.log 2>&1
echo "racecheck rc=$? $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r2ag/pytest_under_racecheck.log; tail -3 gpurun_out/r2ag/racecheck.log; grep -m 8 -B1 -A6 "hazard" gpurun_out/r2ag/racecheck.log | cut -c1-260 | head -60
T0=$(date +%s)
timeout 150 compute-sanitizer --tool synccheck --kernel-name kns=3l2b --print-limit 30 --log-file gpurun_out/r2ag/synccheck.log \
  python -m pytest tests/test_gpu_u1.py tests/test_gpu_vnet.py tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_su3.py -q -m gpu -p no:cacheprovider > gpurun_out/r2ag/pytest_under_synccheck.log 2>&1
echo "synccheck rc=$? $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r2ag/pytest_under_synccheck.log; tail -3 gpurun_out/r2ag/synccheck.log
