# round 2, call am: the GPU suite at the round's final library state
mkdir -p gpurun_out/r2am
timeout 100 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2am/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2am/pytest_gpu.log
