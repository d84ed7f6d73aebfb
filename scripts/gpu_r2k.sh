#!/bin/bash
# round-2 session K: free-surface rows as the top-chunk launch of the interior kernel (TOPK instantiations; the kernels of the other
# chunks are the round's old code): parity suite, bench with / without the extra stream, against the separate k_top launch;
# per-launch times of the timed region; ncu launch list
OUT=gpurun_out/r2k
mkdir -p $OUT
echo "== pytest gpu (fused free surface, TOPK launch)"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" CGFD_PROFILE_DUMP=$OUT/launch_ms_$name.txt timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run fused iso A=1
run fused_1stream iso CGFD_TOP_STREAM=0
run unfused iso CGFD_FUSE_TOP=0
run fused_vti vti A=1
run fused_aniso aniso A=1
run fused_visco visco A=1
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
ls $OUT
