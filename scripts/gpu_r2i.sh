#!/bin/bash
# round-2 session I: where the CFS-PML planes lose their time (aux-record load latency?) -- L1 staging of the records, carve-out
OUT=gpurun_out/r2i
mkdir -p $OUT
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run base iso A=1
run noauxload iso CGFD_L2MODE=515
run carve86 iso CGFD_CARVE=86
run l1pf iso CGFD_L2MODE=1027
run l1pf_carve86 iso CGFD_L2MODE=1027 CGFD_CARVE=86
run l1pf_noalloc_carve86 iso CGFD_L2MODE=3075 CGFD_CARVE=86
run noalloc_carve86 iso CGFD_L2MODE=2051 CGFD_CARVE=86
run l1pf_noalloc_carve86_vti vti CGFD_L2MODE=3075 CGFD_CARVE=86
ls $OUT
