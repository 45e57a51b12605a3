#!/bin/bash
# bench lines + launch list only (no soak, no full ncu capture)
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_$tag.log 2>&1 && tail -1 $out/bench_$tag.log > $out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.log 2>&1 && tail -1 $out/bench_ref_$tag.log > $out/bench_ref_$tag.json
python bench.py --steps 120 --warmup 60 --no-cpu > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 120 --warmup 60 --no-cpu > $out/ncu1_$tag.log 2>&1
