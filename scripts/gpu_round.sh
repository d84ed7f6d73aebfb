#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list and a full capture of the top kernel.
# usage: scripts/gpu_round.sh <tag> [skip_ncu]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ -z "$2" ]; then
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 16 -c 3 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "rc=$?"
fi
ls -la $OUT
