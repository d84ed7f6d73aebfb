#!/bin/bash
# round-2 session R: chunk length around the plan's 25 rows at 400x400x200 (800x800x400 preferred 20 over 25 by 0.7 %), per medium
OUT=gpurun_out/r2r
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'incl pml',d['roofline']['frac_incl_pml_aux'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
SZ=""
run iso_z22 iso A=1
run iso_z20 iso CGFD_ZCHUNK=20
run iso_z18 iso CGFD_ZCHUNK=18
run iso_z16 iso CGFD_ZCHUNK=16
run iso_z25 iso CGFD_ZCHUNK=25
run visco_z20 visco CGFD_ZCHUNK=20
run aniso_z20 aniso CGFD_ZCHUNK=20
run vti_z20 vti CGFD_ZCHUNK=20
SZ="--size 800x800x400"
run big_z16 iso CGFD_ZCHUNK=16
run big_z18 iso CGFD_ZCHUNK=18
ls $OUT | head -3
