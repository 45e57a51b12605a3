#!/bin/bash
# one 512-thread block per SM (128 or 112 envs) against two 256-thread blocks
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for rep in 1 2; do
QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 2v2 16384 300 1500 2>&1 | tail -1
MSV_LIB=$D/libmasurv_t512.so QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 2v2 16384 300 1500 2>&1 | tail -1
MSV_EPB=112 MSV_LIB=$D/libmasurv_t512.so QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 2v2 16384 300 1500 2>&1 | tail -1
MSV_EPB=120 MSV_LIB=$D/libmasurv_t512.so QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 2v2 16384 300 1500 2>&1 | tail -1
MSV_LIB=$D/libmasurv_t512.so QB_ROT=2 timeout 300 python tests/gpu_quickbench.py ffa 8192 300 1500 2>&1 | tail -1
MSV_LIB=$D/libmasurv_t512.so QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 1v1_heal_only 4096 300 1500 2>&1 | tail -1
done | tee gpurun_out/r02z5_ab.txt
