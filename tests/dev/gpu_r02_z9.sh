#!/bin/bash
# phase barriers, one dropped at a time (one 512-thread block per SM)
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for lib in libmasurv.so libmasurv_m3FE.so libmasurv_m3FD.so libmasurv_m3FB.so libmasurv_m3F7.so libmasurv_m3EF.so libmasurv_m3DF.so libmasurv_m3BF.so libmasurv_m37F.so libmasurv_m2FF.so libmasurv.so; do
for v in "2v2 16384 3" "ffa 8192 2"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done | tee gpurun_out/r02z9_ab.txt
