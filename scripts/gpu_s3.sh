#!/bin/bash
OUT=gpurun_out/s3; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== pytest ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "Trainer.forward|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | tail -14
echo "=== path divergence / sdf error per mode ($(date +%T))"
for cfg in "fp32::" "tc2acc::" "tc:0:0" "tc:1:0" "tc:1:4" "tc:1:8" "tc:1:12" "tc:0:12" "tc:0:24" "tc:0:36"; do
  IFS=: read prec order deb <<< "$cfg"
  NA_TM_ORDER=${order:-1} NA_TM_DEBIAS=${deb:-0} timeout 300 python scripts/path_div.py $prec 2>/dev/null | tail -1 | tee -a $OUT/path_div.jsonl
done
echo "=== speed order 0 vs 1 ($(date +%T))"
for order in 0 1; do
  NA_TM_ORDER=$order NA_CHECK_MODES=tc timeout 300 python scripts/tc_check.py > $OUT/tc_check_order$order.log 2>&1; grep -E "^tc |CTA0" $OUT/tc_check_order$order.log
done
echo "=== bench ($(date +%T))"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 2500 $OUT/bench.json; tail -3 $OUT/bench.err
NA_TM_ORDER=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_order0.json 2> $OUT/bench_order0.err; python -c "import json; d=json.load(open('$OUT/bench_order0.json')); print('order0 ms', d['ms_per_step'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
echo "=== done ($(date +%T))"
