#!/bin/bash
mkdir -p gpurun_out
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02t_prof_2v2.txt 2>&1
QB_ROT=2 timeout 300 python tests/gpu_quickbench.py --prof ffa 8192 > gpurun_out/r02t_prof_ffa.txt 2>&1
cat gpurun_out/r02t_prof_2v2.txt gpurun_out/r02t_prof_ffa.txt | grep -v "slowest\|leader-lane\|barrier waits"
