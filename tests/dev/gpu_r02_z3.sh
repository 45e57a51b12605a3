#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for off in 1 0; do
for v in "2v2 16384 3" "ffa 8192 2" "1v1_heal_only 4096 3"; do set -- $v; MSV_NO_HANDOFF=$off QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02z3_ab.txt
