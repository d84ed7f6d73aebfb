#!/bin/bash
# side measurements: the other constitutive laws at 400x400x200 and the isotropic case at 800x800x400, one GPU
OUT=gpurun_out/${1:-media}
mkdir -p $OUT
for M in vti aniso visco; do
  echo "== bench --medium $M"
  timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --medium $M > $OUT/bench_$M.json 2> $OUT/bench_$M.err
  python -c "
import json
d=json.load(open('$OUT/bench_$M.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'finite',d['finite'])
" || tail -3 $OUT/bench_$M.err
done
echo "== bench iso 800x800x400"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --size 800x800x400 > $OUT/bench_iso800.json 2> $OUT/bench_iso800.err
python -c "
import json
d=json.load(open('$OUT/bench_iso800.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'main avg ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],'e2e',d['e2e']['value'],'finite',d['finite'])
" || tail -3 $OUT/bench_iso800.err
