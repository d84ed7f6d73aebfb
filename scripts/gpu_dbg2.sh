#!/bin/bash
OUT=gpurun_out/${1:-dbg2}; mkdir -p $OUT
for cfg in "800x800x400 1 0" "800x800x400 1 49" "800x800x100 1 0" "400x400x400 1 0" "800x400x200 1 0"; do
  set -- $cfg
  echo "== $1 steps=$2 zchunk=$3"; CGFD_ZCHUNK=$3 timeout 300 python scripts/debug_big2.py $1 $2 2>&1 | tail -18
done
