mkdir -p gpurun_out/r2x
for i in 1 2; do timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline --cuda-graphs 2>gpurun_out/r2x/train_$i.err | grep '^{' | cut -c1-200; done
timeout 600 python bench.py --workload su3_8x8x8x8_nb32_l2hmc_train_bf16 --no-cpu-baseline 2>/dev/null | grep '^{' | cut -c1-200
python profiles/prof_l2hmc.py train 8 32 4 256 3 --graph 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,utilization.gpu --format=csv
