#!/bin/bash
# does the programmatic dependent launch need a non-legacy stream?
mkdir -p gpurun_out
for rep in 1 2; do
for strm in 0 1; do for off in 1 0; do
for v in "2v2 16384 3" "ffa 8192 2"; do set -- $v; QB_STREAM=$strm MSV_NO_HANDOFF=$off QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done; done | tee gpurun_out/r02z2_ab.txt
