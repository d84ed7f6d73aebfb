#!/bin/bash
# round-2 session S (4 GPUs, 2x2 grid: every rank has an x and a y neighbour): bench --gpus 4 with its N-rank value check, the
# fused free surface in boundary and interior phase, the chunk rule of r2p / r2r
OUT=gpurun_out/r2s
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== bench --gpus 4"; S0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > $OUT/bench_n4.json 2> $OUT/bench_n4.err
echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_n4.json; tail -3 $OUT/bench_n4.err
