mkdir -p gpurun_out/r2a
export L2B_TEST_REUSE_FORCE=1 L2B_RECT_KERNEL_AUTOGRAD=1
timeout 900 python -m pytest tests -m gpu -q -rs --durations=10 > gpurun_out/r2a/pytest_optin.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r2a/pytest_optin.log
unset L2B_TEST_REUSE_FORCE L2B_RECT_KERNEL_AUTOGRAD
timeout 600 python profiles/time_reference_gpu.py > gpurun_out/r2a/ref_gpu.log 2>&1; echo "refgpu rc=$?"; grep '^{' gpurun_out/r2a/ref_gpu.log
timeout 600 python bench.py > gpurun_out/r2a/bench_hot.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2a/bench_hot.log | cut -c1-400
timeout 600 python bench.py --thermalise 100 --no-cpu-baseline > gpurun_out/r2a/bench_therm.log 2>&1; echo "bench therm rc=$?"; grep '^{' gpurun_out/r2a/bench_therm.log | cut -c1-300
for w in su3_8x8x8x8_nb256_l2hmc_eval_bf16 su3_8x8x8x8_nb32_l2hmc_train_bf16 su3_8x8x8x8_nb256_nlf10_c128 u1_64x64_nb4096_nlf10_f32; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2a/bench_$w.log 2>&1; echo "$w rc=$?"; grep '^{' gpurun_out/r2a/bench_$w.log | cut -c1-300
done
nvidia-smi topo -m > gpurun_out/r2a/topo.txt 2>&1; lscpu | head -30 > gpurun_out/r2a/lscpu.txt; numactl -H >> gpurun_out/r2a/lscpu.txt 2>&1
