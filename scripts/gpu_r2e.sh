#!/bin/bash
# round-2 session E (2 GPUs): N-rank parity tests, 2-rank drop-in against the 2-rank reference, bench --gpus 2, synccheck of the named barrier
OUT=gpurun_out/r2e
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== synccheck visco / aniso (named barrier)"
for MED in visco aniso; do timeout 500 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize_case.py $MED 2 > $OUT/synccheck_$MED.log 2>&1; grep -E "ERROR SUMMARY|sanitize_case" $OUT/synccheck_$MED.log; done
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log; tail -30 $OUT/pytest_multi.log
echo "== pytest media (named barrier)"; timeout 900 python -m pytest tests/test_gpu_media.py -q -m gpu > $OUT/pytest_media.log 2>&1; tail -3 $OUT/pytest_media.log
echo "== bench --gpus 2"; S0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
echo "== bench --gpus 2 visco 600x600x300 global (strong)"; S0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 16 --warmup 3 --medium visco --global-size 600x600x300 --short-e2e > $OUT/bench_n2_visco.json 2> $OUT/bench_n2_visco.err
echo "rc=$? wall $(( $(date +%s) - S0 )) s"; cat $OUT/bench_n2_visco.json; tail -5 $OUT/bench_n2_visco.err
ls $OUT
