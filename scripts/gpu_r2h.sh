#!/bin/bash
# round-2 session H: cost attribution of the 400x400x200 step (which boundary costs what): PML off / single axes / free surface off
OUT=gpurun_out/r2h
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run base iso A=1
run nopml iso BENCH_DIAG=pml=none
run nofree iso BENCH_DIAG=free=0
run nopml_nofree iso BENCH_DIAG=pml=none,free=0
run pmlx iso BENCH_DIAG=pml=x
run pmly iso BENCH_DIAG=pml=y
run pmlz iso BENCH_DIAG=pml=z
ls $OUT
