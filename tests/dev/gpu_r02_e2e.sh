#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "step_host or async or lifetime or tensor_outlives or hand_off" ) 2>&1 | tail -5
timeout 600 python bench.py --workload 2v2 --no-cpu --no-phase > gpurun_out/bench_e2e_2v2.json 2>gpurun_out/bench_e2e.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_e2e_2v2.json').read().strip().splitlines()[-1]); print('%.4e'%d['value'], round(d['ms_per_step'],4), 'e2e %.4e'%d['e2e']['value'], d['e2e']['ms_per_step'], 'sync', d['e2e']['sync']['ms_per_step'], 'obs', d['e2e_obs']['ms_per_step'], d['e2e_obs']['pipelined']['ms_per_step'])"
