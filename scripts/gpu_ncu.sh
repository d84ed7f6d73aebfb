#!/bin/bash
# ncu launch list + one full capture of the main kernel, exported to CSV on the box (the .ncu-rep with sources is too
# large to travel back when several launches are captured). usage: scripts/gpu_ncu.sh <tag> [bench args...]
TAG=${1:-ncu}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
nproc > $OUT/nproc.txt
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 16 -c 4 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_main.ncu-rep --page raw --csv > $OUT/prof_main_raw.csv 2>/dev/null
ncu -i $OUT/prof_main.ncu-rep --page source --csv --print-source sass > $OUT/prof_main_sass.csv 2>/dev/null
ncu -i $OUT/prof_main.ncu-rep --page details > $OUT/prof_main_details.txt 2>/dev/null
echo "== ncu top"; timeout 600 ncu --set full --clock-control none -k regex:k_top -s 5 -c 1 -o $OUT/prof_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_top.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_top.ncu-rep --page raw --csv > $OUT/prof_top_raw.csv 2>/dev/null
# the reports themselves do not travel (gpurun_out is capped at 64 MiB): everything needed was exported to CSV above
rm -f $OUT/*.ncu-rep
ls -la $OUT
