#!/bin/bash
# bench lines of the final build (>= 3 000 timed steps per arm, concurrent_groups arm), reference arm, driver-style flags
tag=r02h
out=gpurun_out
mkdir -p $out
for w in 2v2 ffa ffa_lidar 1v1_heal_only; do
  timeout 900 python bench.py --workload $w > $out/bench_${tag}_$w.json 2>> $out/${tag}.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_${tag}_2v2.json 2>> $out/${tag}.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_${tag}_driver_defaults.json 2>> $out/${tag}.err; echo "bench driver-style rc=$?"
for f in $out/bench_${tag}_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); g=d.get('concurrent_groups',{}); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), 'e2e %.3e'%d['e2e']['value'], 'groups %.3e / %.3e'%(g.get('value',0), g.get('e2e',{}).get('value',0)), d['timing']['repeats'])"; done
