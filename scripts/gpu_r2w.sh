#!/bin/bash
# round-2 session W: arrive / wait split of the per-plane barrier (iso, VTI): parity suite, racecheck + synccheck, bench
OUT=gpurun_out/r2w
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'incl pml',d['roofline']['frac_incl_pml_aux'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
SZ=""
run iso iso A=1
run iso_b iso A=1
run vti vti A=1
run iso_nopml iso BENCH_DIAG=pml=none
SZ="--size 800x800x400"
run big iso A=1
san() { local tool=$1 name=$2 med=$3 nt=$4; shift 4; env "$@" timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py $med $nt > $OUT/${tool}_$name.log 2>&1; echo "$tool $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_$name.log | tail -1) | $(grep sanitize_case $OUT/${tool}_$name.log | tail -1)"; }
san racecheck iso iso 2 A=1
san synccheck iso iso 2 A=1
san racecheck vti vti 2 A=1
