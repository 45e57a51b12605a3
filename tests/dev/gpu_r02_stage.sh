#!/bin/bash
# k_obs2 stage rows padded to an odd stride (no bank conflicts): tile latency and whole step, before/after
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in "2v2 16384" "ffa 8192"; do set -- $v; MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --trace $1 $2 2>&1 | tail -1 | cut -c1-520; done
for rep in 1 2; do for lib in libmasurv_prev.so libmasurv.so; do
for v in "2v2 16384 3" "ffa 8192 2" "ffa_lidar 32768 1"; do set -- $v; MSV_LIB=$D/$lib QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1 | sed 's/, [0-9.e+]* env-steps\/s, \([0-9.e+]* agent-steps\/s\).*/ \1/'; done
done; done | tee gpurun_out/r02_stage_ab.txt
