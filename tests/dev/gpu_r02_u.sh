#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
MSV_LIB=$P timeout 300 python tests/gpu_quickbench.py --blocks ffa 8192 > gpurun_out/r02u_blocks_ffa.txt 2>&1
MSV_LIB=$P timeout 300 python tests/gpu_quickbench.py --blocks 2v2 16384 > gpurun_out/r02u_blocks_2v2.txt 2>&1
cat gpurun_out/r02u_blocks_ffa.txt gpurun_out/r02u_blocks_2v2.txt
for e in 32 24 16 8; do MSV_EPB=$e QB_ROT=2 timeout 300 python tests/gpu_quickbench.py ffa 8192 400 1500 2>&1 | tail -1; done | tee gpurun_out/r02u_epb_ffa.txt
