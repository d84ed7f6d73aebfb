#!/bin/bash
# round-2 session O: ZA tiles at 800x800x400 (A/B against a -DCGFD_NO_ZA build of the same source), the tests of the non-default
# free-surface routes
OUT=gpurun_out/r2o
mkdir -p $OUT
echo "== pytest free-surface routes"; timeout 900 python -m pytest tests/test_gpu_free_surface_routes.py -q -m gpu > $OUT/pytest_routes.log 2>&1; echo "rc=$?" >> $OUT/pytest_routes.log; tail -8 $OUT/pytest_routes.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
NOZA=$PWD/cgfd3d_b200/variants/lib_noza.so
SZ="--size 800x800x400"
run big_za_unfused iso CGFD_FUSE_TOP=0
run big_noza_unfused iso CGFD_FUSE_TOP=0 CGFD_LIB=$NOZA
run big_za_fused24 iso CGFD_TOP_ROWS=24
run big_noza_fused24 iso CGFD_TOP_ROWS=24 CGFD_LIB=$NOZA
SZ=""
run noza_unfused iso CGFD_FUSE_TOP=0 CGFD_LIB=$NOZA
run za_fused24 iso CGFD_TOP_ROWS=24
ls $OUT
