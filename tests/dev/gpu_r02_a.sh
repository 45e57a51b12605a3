#!/bin/bash
# round-2 first GPU pass: parity suite, stationary bench (both arms), per-phase profile of k_step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02a_bench_20.json 2>> gpurun_out/r02a_bench.err; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02a_bench_ref.json 2>> gpurun_out/r02a_bench.err; echo "ref rc=$?"
MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so timeout 300 python tests/gpu_quickbench.py --prof > gpurun_out/r02a_prof.txt 2>&1; echo "prof rc=$?"
timeout 300 python tests/gpu_quickbench.py --sweep > gpurun_out/r02a_sweep.txt 2>&1
tail -3 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_prof.txt; head -c 1500 gpurun_out/r02a_bench.json
