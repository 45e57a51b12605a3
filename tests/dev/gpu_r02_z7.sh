#!/bin/bash
# next step's k_step launched programmatically dependent on the observation kernel: gap between steps, A/B
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
for pdl in 0 1; do for ns in 0 1; do
echo "STEP_PDL=$pdl NO_SPARE=$ns"; MSV_STEP_PDL=$pdl MSV_NO_SPARE=$ns MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --trace 2v2 16384 2>&1 | tail -2
done; done | tee gpurun_out/r02z7_trace.txt
for rep in 1 2; do for pdl in 0 1; do
for v in "2v2 16384 3" "ffa 8192 2" "ffa_lidar 32768 1"; do set -- $v; echo -n "STEP_PDL=$pdl "; MSV_STEP_PDL=$pdl QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02z7_ab.txt
