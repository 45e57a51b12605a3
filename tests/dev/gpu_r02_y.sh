#!/bin/bash
# session 3 baseline: GPU tests + quick timings + block timeline of HEAD
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02y_pytest.log 2>&1; tail -4 gpurun_out/r02y_pytest.log
for v in "2v2 16384 3" "ffa 8192 2" "1v1_heal_only 4096 3"; do set -- $v; QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 400 1500 2>&1 | tail -1; done | tee gpurun_out/r02y_qb.txt
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --blocks 2v2 16384 > gpurun_out/r02y_blocks_2v2.txt 2>&1
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 300 1500 > gpurun_out/r02y_prof_2v2.txt 2>&1
cat gpurun_out/r02y_blocks_2v2.txt gpurun_out/r02y_prof_2v2.txt
