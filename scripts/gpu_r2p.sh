#!/bin/bash
# round-2 session P: z-chunk length at 800x800x400 (the plan picks 198-row chunks there: 2500 tiles x 2 chunks already make 16 waves),
# with and without ZA tiles
OUT=gpurun_out/r2p
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
NOZA=$PWD/cgfd3d_b200/variants/lib_noza.so
SZ="--size 800x800x400"
run big_za_fused24_z25 iso CGFD_TOP_ROWS=24 CGFD_ZCHUNK=25
run big_za_fused24_z50 iso CGFD_TOP_ROWS=24 CGFD_ZCHUNK=50
run big_za_unfused_z25 iso CGFD_FUSE_TOP=0 CGFD_ZCHUNK=25
run big_noza_unfused_z25 iso CGFD_FUSE_TOP=0 CGFD_ZCHUNK=25 CGFD_LIB=$NOZA
run big_noza_unfused_z50 iso CGFD_FUSE_TOP=0 CGFD_ZCHUNK=50 CGFD_LIB=$NOZA
run big_noza_fused24_z25 iso CGFD_TOP_ROWS=24 CGFD_ZCHUNK=25 CGFD_LIB=$NOZA
ls $OUT
