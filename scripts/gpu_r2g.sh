#!/bin/bash
# round-2 session G: PML face masks (parity + bench), launch-plan sweeps for the one-block media, final profile captures
OUT=gpurun_out/r2g
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run iso iso A=1
run iso_b iso A=1
run visco visco A=1
run visco_w32 visco CGFD_WAVES=32
run visco_w8 visco CGFD_WAVES=8
run aniso aniso A=1
run aniso_w32 aniso CGFD_WAVES=32
run vti vti A=1
run iso_w24 iso CGFD_WAVES=24
run iso_w12 iso CGFD_WAVES=12
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full main"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 32 -c 4 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_main.ncu-rep --page raw --csv > $OUT/prof_main_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls $OUT
