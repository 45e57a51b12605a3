#!/bin/bash
# experiment matrix: barrier masks / prefetch / block sizes at the stationary episode mix
mkdir -p gpurun_out; O=gpurun_out/r02c_matrix.txt; : > $O
timeout 900 python -m pytest tests -m gpu -x -q -k "lockstep_rollout or fast_zone or step_host_obs" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
QB="python tests/gpu_quickbench.py 2v2 16384 400 1500"
for rot in 1 4; do
  export QB_ROT=$rot
  $QB >> $O 2>&1
  for v in nopf s000 s01e s2ff s25f s012; do MSV_LIB=$PWD/gpurun_tmp/lib_$v.so $QB >> $O 2>&1; done
done
export QB_ROT=4
for epb in 64 56 40 32 16 8; do MSV_EPB=$epb $QB >> $O 2>&1; done
for epb in 64 32 16 8; do MSV_EPB=$epb MSV_LIB=$PWD/gpurun_tmp/lib_s000.so $QB >> $O 2>&1; done
tail -3 gpurun_out/r02c_pytest.log; grep "us/step" $O
