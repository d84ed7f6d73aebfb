#!/bin/bash
# round-2 last session: rows of the TOPK launch with the 18-row chunks of the final plan
OUT=gpurun_out/r2zz
mkdir -p $OUT
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
for T in 24 32 40; do CGFD_TOP_ROWS=$T timeout 200 python bench.py $B > $OUT/bench_top$T.json 2> $OUT/bench_top$T.err; python -c "
import json; d=json.load(open('$OUT/bench_top$T.json')); print('top rows $T', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'])"; done
