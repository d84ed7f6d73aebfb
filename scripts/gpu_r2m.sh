#!/bin/bash
# round-2 session M (2 GPUs): the fused free surface (TOPK launches in the boundary and the interior phase) under a halo exchange:
# N ranks = 1 rank, 2-rank drop-in against the 2-rank reference, bench --gpus 2 with its N-rank value check
OUT=gpurun_out/r2m
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== pytest multi"; timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log; tail -5 $OUT/pytest_multi.log
echo "== bench --gpus 2"; S0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
ls $OUT
