#!/bin/bash
# bench sweeps over several environment switches in one session. usage: scripts/gpu_sweep.sh <tag> "VAR=v1,v2,..." ...
TAG=${1:-sweep}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$RUN_TESTS" ]; then echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log; fi
for SPEC in "$@"; do
  VAR=${SPEC%%=*}; VALS=${SPEC#*=}
  for V in ${VALS//,/ }; do
    echo "== bench $VAR=$V"
    env $VAR=$V timeout 600 python bench.py --steps ${STEPS:-40} --warmup 3 --no-cpu-baseline $BENCH_ARGS > $OUT/bench_${VAR}_$V.json 2> $OUT/bench_${VAR}_$V.err
    python -c "
import json
d=json.load(open('$OUT/bench_${VAR}_$V.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $OUT/bench_${VAR}_$V.err
  done
done
