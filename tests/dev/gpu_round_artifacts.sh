#!/bin/bash
# Regenerates the per-round evidence under gpurun_out/ (copied into profiles/ afterwards).
# usage (on the GPU box, from the repo root): bash tests/gpu_round_artifacts.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_$tag.log 2>&1 && tail -1 $out/bench_$tag.log > $out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.log 2>&1 && tail -1 $out/bench_ref_$tag.log > $out/bench_ref_$tag.json
python bench.py --steps 120 --warmup 60 --no-cpu > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 120 --warmup 60 --no-cpu > $out/ncu1_$tag.log 2>&1
python bench.py --steps 120 --warmup 60 --no-cpu --rotate 1 > $out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 150 -c 4 -o $out/prof_$tag -f \
    python bench.py --steps 120 --warmup 60 --no-cpu --rotate 1 > $out/ncu2_$tag.log 2>&1
tail -2 $out/ncu2_$tag.log
