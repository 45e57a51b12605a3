#!/bin/bash
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for v in "2v2 16384" "ffa 8192"; do set -- $v
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --trace $1 $2 2>&1 | tail -12
MSV_OBS_CARVEOUT=0 MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --trace $1 $2 2>&1 | tail -4
done | tee gpurun_out/r02z4_trace.txt
for off in 1 0; do for cv in 0 1; do
for v in "2v2 16384 3" "ffa 8192 2"; do set -- $v; echo -n "carveout=$cv "; MSV_OBS_CARVEOUT=$cv MSV_NO_HANDOFF=$off QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02z4_ab.txt
