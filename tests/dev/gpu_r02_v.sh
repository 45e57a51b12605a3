#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "2v2 16384 4" "ffa 8192 2"; do set -- $v; QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 400 1500 2>&1 | tail -1; done | tee gpurun_out/r02v_qb.txt
P=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
MSV_LIB=$P timeout 300 python tests/gpu_quickbench.py --blocks ffa 8192 > gpurun_out/r02v_blocks_ffa.txt 2>&1
MSV_LIB=$P timeout 300 python tests/gpu_quickbench.py --blocks 2v2 16384 > gpurun_out/r02v_blocks_2v2.txt 2>&1
cat gpurun_out/r02v_blocks_ffa.txt gpurun_out/r02v_blocks_2v2.txt
