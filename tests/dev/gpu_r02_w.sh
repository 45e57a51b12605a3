#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
MSV_LIB=$P timeout 300 python tests/gpu_quickbench.py --blocks ffa 8192 > gpurun_out/r02w_blocks_ffa.txt 2>&1
tail -20 gpurun_out/r02w_blocks_ffa.txt
