#!/bin/bash
# round-2 session J: free-surface rows fused into the interior kernel (parity suite + bench, against CGFD_FUSE_TOP=0);
# cost of the PML copy of the plane function without any face (pml=none + l2mode bit 2); ncu source view of one MID launch
OUT=gpurun_out/r2j
mkdir -p $OUT
echo "== pytest gpu (fused free surface)"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -5 ${1%.json}.err; }
B="--steps 24 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base"
run() { local name=$1; shift; local med=$1; shift; env "$@" timeout 600 python bench.py $B --medium $med > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name; }
run fused iso A=1
run unfused iso CGFD_FUSE_TOP=0
run nopml_allpmlcopy iso BENCH_DIAG=pml=none CGFD_L2MODE=7
run nopml iso BENCH_DIAG=pml=none
run fused_vti vti A=1
run fused_aniso aniso A=1
run fused_visco visco A=1
echo "== ncu source view, one MID launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 33 -c 1 -o $OUT/prof_src python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-weak-base > $OUT/ncu_src.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_src.ncu-rep --page source --csv > $OUT/prof_src_source.csv 2>/dev/null
ncu -i $OUT/prof_src.ncu-rep --page raw --csv > $OUT/prof_src_raw.csv 2>/dev/null
ls -la $OUT/prof_src.ncu-rep; rm -f $OUT/*.ncu-rep
ls $OUT
