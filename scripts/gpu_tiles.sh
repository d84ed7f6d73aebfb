#!/bin/bash
# parity tests on the default build, then the bench for the default build and each tile-shape variant in cgfd3d_b200/variants/
# usage: scripts/gpu_tiles.sh <tag> [variant names...]
TAG=${1:-tiles}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$RUN_TESTS" ]; then echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log; fi
show() { python -c "
import json
d=json.load(open('$1'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $2; }
echo "== bench default"; timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench_default.json 2> $OUT/bench_default.err; show $OUT/bench_default.json $OUT/bench_default.err
for V in "$@"; do
  echo "== bench $V"
  CGFD_LIB=$PWD/cgfd3d_b200/variants/lib_$V.so timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench_$V.json 2> $OUT/bench_$V.err; show $OUT/bench_$V.json $OUT/bench_$V.err
done
