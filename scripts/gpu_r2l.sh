#!/bin/bash
# round-2 session L: ZA tiles (the zeta stencil's plane ahead staged by TMA; three-deep register queue) for iso / aniso: parity
# suite on the default (k_top) and on the fused (TOPK) path, bench of both, other media
OUT=gpurun_out/r2l
mkdir -p $OUT
echo "== pytest gpu (default: separate free-surface launch)"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
echo "== pytest gpu subset, CGFD_FUSE_TOP=1"; CGFD_FUSE_TOP=1 timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_boundaries.py tests/test_gpu_media.py -q -m gpu -x > $OUT/pytest_fused.log 2>&1; echo "rc=$?" >> $OUT/pytest_fused.log; tail -4 $OUT/pytest_fused.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" CGFD_PROFILE_DUMP=$OUT/launch_ms_$name.txt timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run iso iso A=1
run iso_fused iso CGFD_FUSE_TOP=1
run aniso aniso A=1
run aniso_fused aniso CGFD_FUSE_TOP=1
run vti vti A=1
run visco visco A=1
run iso_nopml iso BENCH_DIAG=pml=none
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
ls $OUT
