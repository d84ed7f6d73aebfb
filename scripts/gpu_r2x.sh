#!/bin/bash
# round-2 session X: the rebuilt HEAD after the r2w revert: parity suite + one bench line
OUT=gpurun_out/r2x
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json | cut -c1-250
