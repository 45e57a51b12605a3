#!/bin/bash
# phase barriers: combinations around "no barrier between solve, find_new and toi"
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for lib in libmasurv.so libmasurv_m1f7.so libmasurv_m1F3.so libmasurv_m1F5.so libmasurv_m1F1.so libmasurv_m1F6.so libmasurv_m1B7.so libmasurv_m177.so libmasurv_m0F7.so libmasurv_m1E7.so libmasurv_m1D7.so libmasurv_m1f7.so; do
for v in "2v2 16384 3" "ffa 8192 2"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done | tee gpurun_out/r02z10_ab.txt
