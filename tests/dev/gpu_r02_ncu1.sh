#!/bin/bash
mkdir -p gpurun_out
CMD="python tests/gpu_ncu_target.py 2v2 16384 1200 6"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 1200 -c 3 -o gpurun_out/r02_kstep_2v2 -f $CMD > gpurun_out/r02_ncu1.log 2>&1
tail -3 gpurun_out/r02_ncu1.log; ls -la gpurun_out/*.ncu-rep | tail -2
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=2 timeout 300 python tests/gpu_quickbench.py --prof ffa 8192 > gpurun_out/r02r_prof_ffa.txt 2>&1
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02r_prof_2v2.txt 2>&1
cat gpurun_out/r02r_prof_ffa.txt gpurun_out/r02r_prof_2v2.txt | grep -v "slowest"
