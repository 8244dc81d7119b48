#!/bin/bash
# s20: split training program (forward render = forward half, backward launch = GEMMs 21..40) vs the one-launch program
OUT=gpurun_out/s20; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== split test ($(date +%T))"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "split" -s > $OUT/pytest_split.log 2>&1; echo "rc=$?"; grep -E "split vs|passed|failed|Error|error" $OUT/pytest_split.log | head -20
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_dropin_gpu.py -m gpu -q > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_train.log | cut -c1-300
echo "=== bench train ($(date +%T))"
for sp in 1 0; do
NA_BW_SPLIT=$sp timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train_split$sp.json 2> $OUT/bench_train_split$sp.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train_split$sp.json')); print('split $sp', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"; tail -2 $OUT/bench_train_split$sp.err
done
echo "=== done ($(date +%T))"
