# round 2, call ab (re-entry): full GPU suite with the tap-major conv gather / short-tile U(1) heads in the tree, U(1) cfg-1 eval / train steps
mkdir -p gpurun_out/r2ab
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2ab/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ab/pytest_gpu.log
python profiles/prof_u1_l2hmc.py eval 128 5 --graph 2>&1 | tail -1
python profiles/prof_u1_l2hmc.py train 128 5 --graph 2>&1 | tail -1
python profiles/prof_u1_l2hmc.py train 128 5 --table > gpurun_out/r2ab/u1_train_table.txt 2>&1; cut -c1-62,150-240 gpurun_out/r2ab/u1_train_table.txt | sed -n 4,30p
