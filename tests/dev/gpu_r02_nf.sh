#!/bin/bash
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for rep in 1 2 3; do for lib in libmasurv.so libmasurv_nofence.so libmasurv_t448.so; do
for v in "2v2 16384 3"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02_nf_ab.txt
MSV_LIB=$D/libmasurv_t448.so QB_ROT=2 timeout 300 python tests/gpu_quickbench.py ffa 8192 300 1500 2>&1 | tail -1
