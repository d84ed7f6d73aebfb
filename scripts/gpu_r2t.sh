#!/bin/bash
# round-2 session T (8 GPUs, 4x2 grid): configs[2] weak scaling with the round's final code (N-rank value check, weak base, e2e)
OUT=gpurun_out/r2t
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== bench --gpus 8"; S0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_n8.json; tail -3 $OUT/bench_n8.err
