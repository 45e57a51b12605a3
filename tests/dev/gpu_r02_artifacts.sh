#!/bin/bash
# Round-2 evidence: GPU tests, bench lines of the four workloads + the reference arm, the ncu launch list of the
# bench command's timed region, and --set full captures of k_step (2v2, ffa) and k_lidar.  Outputs under gpurun_out/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_PYTEST" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
fi
for w in 2v2 ffa ffa_lidar 1v1_heal_only; do
  timeout 900 python bench.py --workload $w > $out/bench_${tag}_$w.json 2>> $out/${tag}.err; echo "bench $w rc=$?"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_${tag}_2v2.json 2>> $out/${tag}.err; echo "bench ref rc=$?"
# launch list of the bench command's timed region (bench.py brackets it with cudaProfilerStart/Stop)
CMD="python bench.py --steps 60 --warmup 5 --repeats 2 --no-cpu --no-phase"
timeout 600 $CMD > $out/${tag}_plain_launches.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for spec in "2v2 16384 k_step" "ffa 8192 k_step" "ffa_lidar 8192 k_lidar"; do
  set -- $spec
  CMD="python tests/gpu_ncu_target.py $1 $2 1200 6"
  timeout 600 $CMD > $out/${tag}_plain_$1_$3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1200 -c 1 -o $out/${tag}_$3_$1 -f $CMD > $out/${tag}_ncu_$1_$3.log 2>&1
  echo "ncu $1 $3 rc=$?"
  ncu -i $out/${tag}_$3_$1.ncu-rep --page raw --csv > $out/${tag}_$3_$1_raw.csv 2>/dev/null
done
ls -la $out/${tag}_*.ncu-rep
for f in $out/bench_${tag}_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', '%.3e'%d['value'], round(d['ms_per_step'],4), '%.3e'%d['e2e']['value'], d['roofline']['kernel_ms_all'])"; done
du -sh $out   # gpurun copies back at most 64 MiB (the source page of a report is ~100 MB as CSV: export it in the build container)
