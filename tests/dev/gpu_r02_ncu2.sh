#!/bin/bash
# source-level capture of k_step (2v2, one 512-thread block per SM)
mkdir -p gpurun_out
CMD="python tests/gpu_ncu_target.py 2v2 16384 1200 6"
timeout 600 $CMD > gpurun_out/r02b_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 1200 -c 1 -o gpurun_out/r02b_k_step_2v2 -f $CMD > gpurun_out/r02b_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r02b_ncu.log; ls -la gpurun_out/*.ncu-rep
