#!/bin/bash
# round-2 session C: validation of the split free-surface kernel, the staged visco kernel, device dvh2dvz, the new drop-in cases;
# the bench line as the driver runs it (+ reference arm); media benches; sanitizer on the new kernels; ncu of visco / k_top
OUT=gpurun_out/r2c
mkdir -p $OUT
nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],d.get('e2e_k100',{}).get('value'),'weak',d.get('weak_base'),'finite',d['finite'])
if 'cpu_baseline' in d: print('   cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('ranks'), d['cpu_baseline'].get('seconds_per_step'), d['cpu_baseline'].get('setup_s'))
" || tail -5 ${1%.json}.err; }
echo "== bench default"; /usr/bin/time -v timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; grep -E "Elapsed|Maximum resident" $OUT/bench.err; show $OUT/bench.json default
echo "== bench reference arm"; /usr/bin/time -v timeout 1200 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; grep -E "Elapsed|Maximum resident" $OUT/bench_ref.err; cat $OUT/bench_ref.json
for M in visco vti aniso; do
  echo "== bench --medium $M"; timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --short-e2e --medium $M > $OUT/bench_$M.json 2> $OUT/bench_$M.err; show $OUT/bench_$M.json $M
done
echo "== visco unstaged for comparison"; CGFD_VIS_STAGED=0 timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --short-e2e --medium visco > $OUT/bench_visco_unstaged.json 2> $OUT/bench_visco_unstaged.err; show $OUT/bench_visco_unstaged.json visco_unstaged
for MED in visco iso; do
  echo "== memcheck $MED"; timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_case.py $MED 4 > $OUT/memcheck_$MED.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|sanitize_case" $OUT/memcheck_$MED.log
done
echo "== racecheck visco"; timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py visco 2 > $OUT/racecheck_visco.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|sanitize_case" $OUT/racecheck_visco.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full main"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 32 -c 4 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_main.ncu-rep --page raw --csv > $OUT/prof_main_raw.csv 2>/dev/null
echo "== ncu full visco"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 32 -c 4 -o $OUT/prof_vis python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --medium visco > $OUT/ncu_vis.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_vis.ncu-rep --page raw --csv > $OUT/prof_vis_raw.csv 2>/dev/null
ncu -i $OUT/prof_vis.ncu-rep --page source --csv --print-source sass > $OUT/prof_vis_sass.csv 2>/dev/null
echo "== ncu top"; timeout 600 ncu --set full --clock-control none -k regex:k_top -s 8 -c 4 -o $OUT/prof_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --short-e2e --no-weak-base > $OUT/ncu_top.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_top.ncu-rep --page raw --csv > $OUT/prof_top_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls $OUT
