#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log
tail -4 gpurun_out/r02s_pytest.log
for w in 2v2 ffa ffa_lidar 1v1_heal_only; do
timeout 600 python bench.py --workload $w --no-cpu --no-phase > gpurun_out/r02s_bench_$w.json 2>> gpurun_out/r02s.err; echo "bench $w rc=$?"
done
export MSV_LIB=$PWD/gym-ma-survival-2d_b200/masurvival/libmasurv_prof.so
QB_ROT=4 timeout 300 python tests/gpu_quickbench.py --prof 2v2 16384 > gpurun_out/r02s_prof_2v2.txt 2>&1
cat gpurun_out/r02s_prof_2v2.txt
for f in gpurun_out/r02s_bench*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), '%.3e'%d['e2e']['value'], d['roofline']['kernel_ms_all'], d['roofline']['kernel_gbs_all'])"; done
