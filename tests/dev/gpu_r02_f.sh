#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02f_bench_20.json 2>> gpurun_out/r02f_bench.err; echo "bench20 rc=$?"
timeout 900 python bench.py --steps 5000 --warmup 100 --no-cpu --repeats 5 > gpurun_out/r02f_bench_5000.json 2>> gpurun_out/r02f_bench.err; echo "bench5000 rc=$?"
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02f_prof_2v2.txt 2>&1
QB_ROT=2 timeout 300 python tests/gpu_quickbench.py --prof ffa 8192 > gpurun_out/r02f_prof_ffa.txt 2>&1
cat gpurun_out/r02f_prof_*.txt; for f in gpurun_out/r02f_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_obs']['value'], d['e2e_obs']['pipelined']['value'], d['roofline']['kernel_ms_all'], d['timing']['mean_episode_steps_in_timed_region'])"; done
