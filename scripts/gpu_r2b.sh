#!/bin/bash
# round-2 session B: new parity tests + env-switch experiments on the interior kernel (DRAM over-fetch, late free-surface kernel)
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== pytest gpu (new boundary tests first)"; timeout 900 python -m pytest tests/test_gpu_boundaries.py -q -m gpu > $OUT/pytest_bnd.log 2>&1; echo "rc=$?" >> $OUT/pytest_bnd.log; tail -15 $OUT/pytest_bnd.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_boundaries.py --deselect "tests/test_gpu_dropin.py::test_dropin_matches_golden[hill100]" > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
show() { python -c "
import json
d=json.load(open('$1'))
print('$2 value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -3 ${1%.json}.err; }
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_data_ecc.sum"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 24 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; show $OUT/bench_$name.json $name
}
traffic() {
  local name=$1; shift
  env "$@" timeout 600 ncu --metrics $M --clock-control none -k regex:k_main_tma -s 32 -c 4 --csv --log-file $OUT/traffic_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/traffic_$name.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/traffic_$name.csv')) if len(r)>10]
h=rows[0]; iK=h.index('Kernel Name'); iM=h.index('Metric Name'); iV=h.index('Metric Value'); iI=h.index('ID')
d={}
for r in rows[1:]: d.setdefault(r[iI],{})[r[iM]]=float(r[iV].replace(',',''))
tr=sum(m['dram__bytes_read.sum'] for m in d.values())/1e9; tw=sum(m['dram__bytes_write.sum'] for m in d.values())/1e9
print('$name traffic per step: read %.3f GB write %.3f GB total %.3f GB; ecc sectors %.1fM; per launch read'%(tr,tw,tr+tw,sum(m['lts__t_sectors_data_ecc.sum'] for m in d.values())/1e6), ['%.2f'%(m['dram__bytes_read.sum']/1e9) for m in d.values()])
PY
}
run base A=1
traffic base A=1
run promo_cur0 CGFD_L2PROMO_CUR=0
traffic promo_cur0 CGFD_L2PROMO_CUR=0
run promo_cur1 CGFD_L2PROMO_CUR=1
traffic promo_cur1 CGFD_L2PROMO_CUR=1
run lpt0 CGFD_LPT=0
traffic lpt0 CGFD_LPT=0
run lpt1 CGFD_LPT=1
run toplate CGFD_TOPPAR=2
run toppar1 CGFD_TOPPAR=1
run toplate_promo0 CGFD_TOPPAR=2 CGFD_L2PROMO_CUR=0
L=$PWD/cgfd3d_b200/variants/lib_tile4.so
run tile4 CGFD_LIB=$L
run tile4_toplate CGFD_LIB=$L CGFD_TOPPAR=2
run tile4_toplate_promo0 CGFD_LIB=$L CGFD_TOPPAR=2 CGFD_L2PROMO_CUR=0
traffic tile4 CGFD_LIB=$L
echo "== visco"
env timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --medium visco > $OUT/bench_visco.json 2> $OUT/bench_visco.err; show $OUT/bench_visco.json visco
L=$PWD/cgfd3d_b200/variants/lib_visprefetch.so
env CGFD_LIB=$L timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --medium visco > $OUT/bench_visco_pf.json 2> $OUT/bench_visco_pf.err; show $OUT/bench_visco_pf.json visco_prefetch
CGFD_LIB=$L timeout 600 python -m pytest tests/test_gpu_media.py -q -m gpu -k visco > $OUT/pytest_vispf.log 2>&1; tail -2 $OUT/pytest_vispf.log
ls $OUT
