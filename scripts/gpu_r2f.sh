#!/bin/bash
# round-2 session F (8 GPUs): weak scaling of configs[2] on the 4x2 grid, configs[3] (VTI + general anisotropic 600x600x300) and
# configs[4] (visco-elastic 1200x1200x600) at their stated sizes
OUT=gpurun_out/r2f
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
run() { # name, args...
  local name=$1; shift
  S0=$(date +%s)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 8 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "== $name rc=$? wall $(( $(date +%s) - S0 )) s"
  python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_$name.json'))
    print('   value',d['value'],'per_gpu',d.get('per_gpu'),'ms/step',d['ms_per_step'],'main frac',d['roofline']['frac'],'whole',d['roofline']['whole_step_frac'],
          'e2e',(d['e2e'] or {}).get('value'),'k100',(d.get('e2e_k100') or {}).get('value'),'parity',d.get('parity_nrank'),'weak',d.get('weak_base'),'eff',d.get('efficiency_vs_weak_base'))
except Exception as e:
    print('   no line:', e); print(open('$OUT/bench_$name.err').read()[-1500:])
PY
}
run iso_weak --steps 20 --warmup 3
run vti_600 --steps 20 --warmup 3 --medium vti --global-size 600x600x300 --short-e2e
run aniso_600 --steps 20 --warmup 3 --medium aniso --global-size 600x600x300 --short-e2e
run visco_1200 --steps 12 --warmup 3 --medium visco --global-size 1200x1200x600 --no-e2e
ls $OUT
