#!/bin/bash
# round-2 session D: split thread groups (visco / aniso), streaming drop-in driver, the bench line as the driver runs it, reference arm
OUT=gpurun_out/r2d
mkdir -p $OUT
nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],d.get('e2e_k100',{}).get('value'),'weak',d.get('weak_base'),'finite',d['finite'])
if 'cpu_baseline' in d: print('   cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('ranks'), d['cpu_baseline'].get('seconds_per_step'), d['cpu_baseline'].get('setup_s'))
" || tail -5 ${1%.json}.err; }
for M in visco aniso; do
  echo "== bench --medium $M"; timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --short-e2e --medium $M > $OUT/bench_$M.json 2> $OUT/bench_$M.err; show $OUT/bench_$M.json $M
  echo "== bench --medium $M (no split)"; CGFD_LIB=$PWD/cgfd3d_b200/variants/lib_nosplit.so timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --short-e2e --medium $M > $OUT/bench_${M}_nosplit.json 2> $OUT/bench_${M}_nosplit.err; show $OUT/bench_${M}_nosplit.json ${M}_nosplit
done
echo "== bench default"; S0=$(date +%s); timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$? wall $(( $(date +%s) - S0 )) s"; show $OUT/bench.json default; tail -3 $OUT/bench.err
echo "== bench reference arm"; S0=$(date +%s); timeout 1200 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
echo "== drop-in on configs[1]"; timeout 1200 python scripts/dropin_config1.py 200 > $OUT/dropin_config1.log 2>&1; tail -3 $OUT/dropin_config1.log
echo "== memcheck visco aniso (split kernels)"
for MED in visco aniso; do timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_case.py $MED 4 > $OUT/memcheck_$MED.log 2>&1; grep -E "ERROR SUMMARY|sanitize_case" $OUT/memcheck_$MED.log; done
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py visco 2 > $OUT/racecheck_visco.log 2>&1; grep -E "RACECHECK SUMMARY|sanitize_case" $OUT/racecheck_visco.log
timeout 500 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_case.py visco 2 > $OUT/synccheck_visco.log 2>&1; grep -E "ERROR SUMMARY|sanitize_case" $OUT/synccheck_visco.log
echo "== ncu full visco"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 32 -c 4 -o $OUT/prof_vis python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --medium visco > $OUT/ncu_vis.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_vis.ncu-rep --page raw --csv > $OUT/prof_vis_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls $OUT
