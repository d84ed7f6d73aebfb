#!/bin/bash
# N-GPU session: N-rank vs 1-rank parity, then the weak-scaling bench with and without the overlapped halo exchange.
# usage: scripts/gpu_multi.sh <tag> <ngpus> [size]
TAG=${1:-multi}; N=${2:-2}; SIZE=${3:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
echo "== multi-GPU parity"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log; tail -4 $OUT/pytest_multi.log
for OV in 1 0; do
  echo "== bench --gpus $N overlap=$OV"
  CGFD_OVERLAP=$OV timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps 20 --warmup 3 ${SIZE:+--size $SIZE} > $OUT/bench_n${N}_ov$OV.json 2> $OUT/bench_n${N}_ov$OV.err
  echo "rc=$?"; tail -c 1800 $OUT/bench_n${N}_ov$OV.json; tail -3 $OUT/bench_n${N}_ov$OV.err
done
echo "== bench --gpus 1 same block size"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --size ${SIZE:-800x800x400} --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; tail -c 1800 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
