#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r02e_matrix.txt; : > $O
QB="python tests/gpu_quickbench.py 2v2 16384 400 1500"
for rot in 1 4; do
  export QB_ROT=$rot
  $QB >> $O 2>&1
  for v in i0x01 i0x03 i0x0b i0x1f; do MSV_LIB=$PWD/gpurun_tmp/lib_$v.so $QB >> $O 2>&1; done
done
MSV_LIB=$PWD/gpurun_tmp/lib_i0x1f.so timeout 600 python -m pytest tests -m gpu -x -q -k "lockstep or scrambled" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -3 gpurun_out/r02e_pytest.log; grep "us/step" $O
