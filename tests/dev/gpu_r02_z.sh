#!/bin/bash
# tile hand-off (k_step -> k_obs2 through the completion queue, programmatic dependent launch): parity + A/B timing
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || exit 1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02z_pytest.log 2>&1; tail -4 gpurun_out/r02z_pytest.log
for rep in 1 2; do
for off in 1 0; do
for v in "2v2 16384 3" "ffa 8192 2" "1v1_heal_only 4096 3" "ffa_lidar 32768 1"; do set -- $v; echo -n "NO_HANDOFF=$off "; MSV_NO_HANDOFF=$off QB_ROT=$3 timeout 300 python tests/gpu_quickbench.py $1 $2 300 1500 2>&1 | tail -1; done
done; done | tee gpurun_out/r02z_ab.txt
