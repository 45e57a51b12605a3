#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-phase > gpurun_out/r02g_bench_20_$i.json 2>> gpurun_out/r02g.err
BENCH_NO_SMI=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-phase > gpurun_out/r02g_bench_20_nosmi_$i.json 2>> gpurun_out/r02g.err
done
for f in gpurun_out/r02g_bench_20*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['timing']['rep_ms'])"; done
