#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
for v in 2v2 ffa; do
MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --prof $v $([ $v = ffa ] && echo 8192 || echo 16384) > gpurun_out/r02b_prof_$v.txt 2>&1
done
tail -3 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_prof_*.txt
