#!/bin/bash
# one 512-thread block per SM + tile hand-off: GPU tests, the four workloads (hand-off on/off), block timeline
mkdir -p gpurun_out
D=$PWD/gym-ma-survival-2d_b200/masurvival
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || exit 1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02z6_pytest.log 2>&1; tail -4 gpurun_out/r02z6_pytest.log
for off in 1 0; do
for v in "2v2 16384 3" "ffa 8192 2" "1v1_heal_only 4096 3" "ffa_lidar 32768 1"; do set -- $v; MSV_NO_HANDOFF=$off QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done | tee gpurun_out/r02z6_ab.txt
for e in 16 64; do MSV_EPB=$e QB_ROT=3 timeout 300 python tests/gpu_quickbench.py 1v1_heal_only 4096 300 1500 2>&1 | tail -1; done | tee -a gpurun_out/r02z6_ab.txt
for e in 48 64; do MSV_EPB=$e QB_ROT=1 timeout 300 python tests/gpu_quickbench.py ffa_lidar 32768 200 1500 2>&1 | tail -1; done | tee -a gpurun_out/r02z6_ab.txt
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --trace 2v2 16384 2>&1 | tail -3 | tee gpurun_out/r02z6_trace.txt
MSV_LIB=$D/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --blocks 2v2 16384 > gpurun_out/r02z6_blocks_2v2.txt 2>&1
head -20 gpurun_out/r02z6_blocks_2v2.txt
