mkdir -p gpurun_out/r2ad
for i in 1 2; do timeout 600 python bench.py --workload su3_8x8x8x8_nb256_l2hmc_eval_bf16 --no-cpu-baseline --cuda-graphs 2>gpurun_out/r2ad/eval_$i.err | grep '^{' | cut -c1-220; done
python profiles/prof_l2hmc.py eval 8 256 4 256 3 --graph 2>&1 | tail -2
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
