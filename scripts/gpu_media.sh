#!/bin/bash
TAG=${1:-media}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest media"; timeout 1200 python -m pytest tests/test_gpu_media.py -q -m gpu > $OUT/pytest_media.log 2>&1; echo "rc=$?" >> $OUT/pytest_media.log; grep -E "passed|failed|^FAILED|^E  .*(bad|Assert|rel)" $OUT/pytest_media.log | cut -c1-700 | head -40
echo "== pytest iso"; timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_dropin.py -x -q -m gpu > $OUT/pytest_iso.log 2>&1; echo "rc=$?" >> $OUT/pytest_iso.log; tail -3 $OUT/pytest_iso.log
echo "== debug big"; timeout 600 python scripts/debug_big.py 800x800x400 > $OUT/debug_big.log 2>&1; tail -14 $OUT/debug_big.log | cut -c1-900
echo "== debug big nofast"; CGFD_L2MODE=7 timeout 600 python scripts/debug_big.py 800x800x400 > $OUT/debug_big_nofast.log 2>&1; tail -4 $OUT/debug_big_nofast.log | cut -c1-900
