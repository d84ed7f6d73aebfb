#!/bin/bash
# round-2 session N: rows of the TOPK launch (fused free surface) at 400x400x200 and at 800x800x400; ZA tiles at 800x800x400;
# the other media after the removal of the r2i switches
OUT=gpurun_out/r2n
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
SZ=""
run iso_top16 iso A=1
run iso_top12 iso CGFD_TOP_ROWS=12
run iso_top8 iso CGFD_TOP_ROWS=8
run iso_top24 iso CGFD_TOP_ROWS=24
run vti vti A=1
run visco visco A=1
SZ="--size 800x800x400"
run big_top16 iso A=1
run big_top8 iso CGFD_TOP_ROWS=8
run big_top32 iso CGFD_TOP_ROWS=32
run big_unfused iso CGFD_FUSE_TOP=0
ls $OUT
