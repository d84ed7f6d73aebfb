#!/bin/bash
# round-2 session U: HEAD as the driver will run it: smoke(), parity suite, the default bench line, the reference arm
OUT=gpurun_out/r2u
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
echo "== bench (default line)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
