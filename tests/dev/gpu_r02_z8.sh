#!/bin/bash
# phase barriers with one 512-thread block per SM: which ones still pay?
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for rep in 1 2; do
for lib in libmasurv.so libmasurv_m1ff.so libmasurv_m1f7.so libmasurv_m25f.so libmasurv_m005.so libmasurv_nosync.so; do
for v in "2v2 16384 3" "ffa 8192 2"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02z8_ab.txt
