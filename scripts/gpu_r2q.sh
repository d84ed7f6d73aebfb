#!/bin/bash
# round-2 session Q: the round's final code: parity suite, the full default bench line (with weak base and CPU arm), the other media,
# chunk length at 800x800x400 around the new rule, ncu launch list + full capture of the eight k_main_tma launches of one step
OUT=gpurun_out/r2q
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== bench (default line)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'incl pml',d['roofline']['frac_incl_pml_aux'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med $SZ > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
SZ=""
run vti vti A=1
run aniso aniso A=1
run visco visco A=1
run iso_ktop iso CGFD_FUSE_TOP=0
run iso_b iso A=1
run iso_nopf iso CGFD_L2MODE=515
SZ="--size 800x800x400"
run big_default iso A=1
run big_z20 iso CGFD_ZCHUNK=20
run big_z32 iso CGFD_ZCHUNK=32
run big_visco visco A=1
SZ=""
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 64 -c 8 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_main.ncu-rep --page raw --csv > $OUT/prof_main_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls $OUT
