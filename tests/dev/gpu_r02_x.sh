#!/bin/bash
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for rep in 1 2; do
for lib in libmasurv_base.so libmasurv_nocoop.so libmasurv.so; do
for v in "2v2 16384 4" "ffa 8192 2"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 400 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02x_ab.txt
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --blocks ffa 8192 > gpurun_out/r02x_blocks_ffa.txt 2>&1
cat gpurun_out/r02x_blocks_ffa.txt
