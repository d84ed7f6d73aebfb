#!/bin/bash
OUT=gpurun_out/${1:-dbg}
mkdir -p $OUT
echo "== plain"; CGFD_VARIANT=0 timeout 120 python -m pytest tests/test_gpu_iso.py -x -q -k "test_onestage_variants and nopml" > $OUT/plain.log 2>&1; tail -5 $OUT/plain.log
echo "== sanitizer"; CGFD_VARIANT=0 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_iso.py -x -q -k "test_onestage_variants and nopml" > $OUT/sanitizer.log 2>&1; grep -v "^$" $OUT/sanitizer.log | head -60
