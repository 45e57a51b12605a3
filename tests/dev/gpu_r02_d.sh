#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
QB="python tests/gpu_quickbench.py 2v2 16384 400 1500"
O=gpurun_out/r02d_matrix.txt; : > $O
QB_ROT=1 $QB >> $O 2>&1; QB_ROT=4 $QB >> $O 2>&1
QB_ROT=4 python tests/gpu_quickbench.py ffa 8192 300 1500 >> $O 2>&1
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=1 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02d_prof_rot1.txt 2>&1
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02d_prof_rot4.txt 2>&1
tail -3 gpurun_out/r02d_pytest.log; cat $O gpurun_out/r02d_prof_rot*.txt
