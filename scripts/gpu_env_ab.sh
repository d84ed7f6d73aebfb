#!/bin/bash
# bench for each value of an environment variable (no tests). usage: scripts/gpu_env_ab.sh <tag> <ENVVAR> <value>...
TAG=${1:-ab}; VAR=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in "$@"; do
  echo "== bench $VAR=$V"
  env $VAR=$V timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline $BENCH_ARGS > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python -c "
import json
d=json.load(open('$OUT/bench_$V.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $OUT/bench_$V.err
done
