#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r02j_matrix.txt; : > $O
timeout 900 python -m pytest tests -m gpu -x -q -k "lockstep or scrambled or golden" > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
export QB_ROT=4
QB="python tests/gpu_quickbench.py 2v2 16384 600 1500"
$QB >> $O 2>&1
for v in s1ff s3fd s3ef s37f s3bf i02 i12; do MSV_LIB=$PWD/gpurun_tmp/lib_$v.so $QB >> $O 2>&1; done
$QB >> $O 2>&1
QB_ROT=2 python tests/gpu_quickbench.py ffa 8192 300 1500 >> $O 2>&1
QB_ROT=2 MSV_LIB=$PWD/gpurun_tmp/lib_i02.so python tests/gpu_quickbench.py ffa 8192 300 1500 >> $O 2>&1
tail -3 gpurun_out/r02j_pytest.log; grep "us/step" $O
