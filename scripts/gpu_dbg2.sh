#!/bin/bash
OUT=gpurun_out/${1:-dbg2}; mkdir -p $OUT
for mode in 3 3 3 3 3 3; do
  echo "== 800x800x400 l2mode=$mode"; CGFD_L2MODE=$mode timeout 600 python scripts/debug_big2.py 800x800x400 2 2>&1 | grep -E "bad planes|k=" | head -12
done
echo "== pytest"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench 400"; python bench.py --steps 40 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2200
echo "== bench 800"; python bench.py --steps 20 --size 800x800x400 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2200
