#!/bin/bash
# A/B of kernel variants on the bench workload + parity tests for each. usage: scripts/gpu_kernel_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in 0; do
  echo "== variant $V: pytest"; CGFD_VARIANT=$V timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_v$V.log 2>&1; echo "rc=$?" >> $OUT/pytest_v$V.log; tail -4 $OUT/pytest_v$V.log
done
for V in 0; do
  for Z in 32 49; do
    echo "== bench variant $V zchunk $Z"
    CGFD_VARIANT=$V CGFD_ZCHUNK=$Z timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_v${V}_z$Z.json 2> $OUT/bench_v${V}_z$Z.err
    python -c "
import json,sys
d=json.load(open('$OUT/bench_v${V}_z$Z.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $OUT/bench_v${V}_z$Z.err
  done
done
echo "== ncu launches (default variant)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iso_main -s 17 -c 2 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "rc=$?"
