#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
for w in 2v2 ffa 1v1_heal_only ffa_lidar; do
timeout 600 python bench.py --workload $w --no-cpu --no-phase > gpurun_out/r02h_bench_$w.json 2>> gpurun_out/r02h.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-phase > gpurun_out/r02h_bench_20.json 2>> gpurun_out/r02h.err
tail -3 gpurun_out/r02h_pytest.log
for f in gpurun_out/r02h_bench*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), '%.3e'%d['e2e']['value'], '%.3e'%d['e2e_obs']['value'], '%.3e'%d['e2e_obs']['pipelined']['value'], d['roofline']['kernel_ms_all'], d['timing']['rep_ms'][:6])"; done
