#!/bin/bash
# round-2 session A: baseline check, DRAM-traffic diagnostics of the interior kernel, pending variants, sanitizers, visco profile
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1; nproc > $OUT/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum"
for L in 3 131 259 387; do
  echo "== traffic CGFD_L2MODE=$L"
  CGFD_L2MODE=$L timeout 600 ncu --metrics $M --clock-control none -k regex:k_main_tma -s 32 -c 8 --csv --log-file $OUT/traffic_$L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/traffic_$L.log 2>&1; echo "rc=$?"
  CGFD_L2MODE=$L timeout 600 python bench.py --steps 24 --warmup 3 --no-cpu-baseline > $OUT/bench_l2mode_$L.json 2> $OUT/bench_l2mode_$L.err
  python -c "
import json
d=json.load(open('$OUT/bench_l2mode_$L.json'))
print('L2MODE=$L value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'])
"
done
echo "== variants"; scripts/gpu_variant_check.sh r2a_var toptiled tile4
for MED in iso vti aniso visco; do
  echo "== memcheck $MED"; timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_case.py $MED 4 > $OUT/memcheck_$MED.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|sanitize_case" $OUT/memcheck_$MED.log
done
for MED in iso; do
  echo "== racecheck $MED"; timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py $MED 2 > $OUT/racecheck_$MED.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|sanitize_case" $OUT/racecheck_$MED.log
  echo "== synccheck $MED"; timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_case.py $MED 2 > $OUT/synccheck_$MED.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|sanitize_case" $OUT/synccheck_$MED.log
done
echo "== ncu full visco"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 16 -c 4 -o $OUT/prof_vis python bench.py --steps 2 --warmup 3 --no-cpu-baseline --medium visco > $OUT/ncu_vis.log 2>&1; echo "rc=$?"
ncu -i $OUT/prof_vis.ncu-rep --page raw --csv > $OUT/prof_vis_raw.csv 2>/dev/null
ncu -i $OUT/prof_vis.ncu-rep --page source --csv --print-source sass > $OUT/prof_vis_sass.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT
