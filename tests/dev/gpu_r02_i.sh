#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
timeout 900 python tests/gpu_soak.py > gpurun_out/r02i_soak.log 2>&1; echo "soak rc=$?"
for w in 2v2 ffa; do
timeout 600 python bench.py --workload $w --no-cpu --no-phase > gpurun_out/r02i_bench_$w.json 2>> gpurun_out/r02i.err; echo "bench $w rc=$?"
done
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02i_prof_2v2.txt 2>&1
QB_ROT=2 timeout 300 python tests/gpu_quickbench.py --prof ffa 8192 > gpurun_out/r02i_prof_ffa.txt 2>&1
tail -3 gpurun_out/r02i_pytest.log; tail -5 gpurun_out/r02i_soak.log; cat gpurun_out/r02i_prof_*.txt
for f in gpurun_out/r02i_bench*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), '%.3e'%d['e2e']['value'], d['roofline']['kernel_ms_all'], d['timing']['rep_ms'][:6])"; done
