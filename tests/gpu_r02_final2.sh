#!/bin/bash
# final refresh: GPU tests (incl. production tile shapes), soak, bench lines (e2e with the upload stream), reference arm
tag=r02g
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
grep -n "passed\|failed\|rc=" $out/${tag}_pytest.log
( cd tests && timeout 1200 python gpu_soak.py ) > $out/${tag}_soak.log 2>&1; echo "soak rc=$?"; tail -7 $out/${tag}_soak.log
for w in 2v2 ffa ffa_lidar 1v1_heal_only; do
  timeout 900 python bench.py --workload $w > $out/bench_${tag}_$w.json 2>> $out/${tag}.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_${tag}_2v2.json 2>> $out/${tag}.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_${tag}_driver_defaults.json 2>> $out/${tag}.err; echo "bench driver-style rc=$?"
for f in $out/bench_${tag}_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel_ms_all'], d.get('tile_plan'))"; done
