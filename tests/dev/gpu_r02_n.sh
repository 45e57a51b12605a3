#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -5 gpurun_out/r02n_pytest.log
for w in ffa_lidar 2v2; do
timeout 600 python bench.py --workload $w --no-cpu --no-phase > gpurun_out/r02n_bench_$w.json 2>> gpurun_out/r02n.err; echo "bench $w rc=$?"
done
for f in gpurun_out/r02n_bench*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), '%.3e'%d['e2e']['value'], d['roofline']['kernel_ms_all'])"; done
