#!/bin/bash
# usage: scripts/gpu_prof.sh <tag> [variant] : pytest (all, no -x), then ncu full on the main kernel (3 launches)
TAG=${1:-prof}; V=${2:-0}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; CGFD_VARIANT=$V timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log; grep -E "passed|failed|FAILED|Error|assert" $OUT/pytest.log | head -30
echo "== ncu full"; CGFD_VARIANT=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_main_tma -s 17 -c 3 -o $OUT/prof_main python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ls -la $OUT
