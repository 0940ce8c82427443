# round 2, call af: compute-sanitizer memcheck over the libl2b kernels (mangled names containing 3l2b) while the GPU tests run
mkdir -p gpurun_out/r2af
T0=$(date +%s)
timeout 540 compute-sanitizer --tool memcheck --kernel-name kns=3l2b --error-exitcode 7 --print-limit 30 --log-file gpurun_out/r2af/memcheck.log \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_gemm.py tests/test_gpu_dense.py tests/test_gpu_u1.py tests/test_gpu_vnet.py tests/test_gpu_su3.py tests/test_gpu_training.py tests/test_gpu_dynamics.py -q -m gpu -p no:cacheprovider > gpurun_out/r2af/pytest_under_memcheck.log 2>&1
echo "memcheck rc=$? $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/r2af/pytest_under_memcheck.log
grep -c "Invalid\|misaligned\|out of bounds" gpurun_out/r2af/memcheck.log; tail -5 gpurun_out/r2af/memcheck.log; grep -m 12 -A12 "=========  *Invalid\|========= Invalid" gpurun_out/r2af/memcheck.log | head -80
