#!/bin/bash
# parity suite + bench on variant builds of the library (scripts/build_variant.sh), next to the default build.
# usage: scripts/gpu_variant_check.sh <tag> <variant> [<variant> ...]      (env BENCH_ARGS: extra bench.py arguments, e.g. --medium visco)
TAG=${1:-variants}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -3 $2; }
echo "== bench default"; timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline $BENCH_ARGS > $OUT/bench_default.json 2> $OUT/bench_default.err; show $OUT/bench_default.json $OUT/bench_default.err
for V in "$@"; do
  L=$PWD/cgfd3d_b200/variants/lib_$V.so
  echo "== $V: parity suite (drop-in tests excluded: that binary links the default library)"
  CGFD_LIB=$L timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_dropin.py > $OUT/pytest_$V.log 2>&1; echo "rc=$?" >> $OUT/pytest_$V.log; tail -4 $OUT/pytest_$V.log
  echo "== $V: bench"
  CGFD_LIB=$L timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline $BENCH_ARGS > $OUT/bench_$V.json 2> $OUT/bench_$V.err; show $OUT/bench_$V.json $OUT/bench_$V.err
done
