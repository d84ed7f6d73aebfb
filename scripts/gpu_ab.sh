#!/bin/bash
# quick A/B on one GPU: parity tests, then bench for each setting of an environment variable.
# usage: scripts/gpu_ab.sh <tag> <ENVVAR> <value> [<value> ...]
TAG=${1:-ab}; VAR=${2:-CGFD_OVERLAP}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
for V in "$@"; do
  echo "== bench $VAR=$V"
  env $VAR=$V timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python -c "
import json
d=json.load(open('$OUT/bench_$V.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $OUT/bench_$V.err
done
