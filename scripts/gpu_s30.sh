#!/bin/bash
# s30: render kernels after the epilogue-context changes (regression check)
OUT=gpurun_out/s30; mkdir -p $OUT
export PYTHONUNBUFFERED=1
NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check.log 2>&1; grep -E "^tc|Linf" $OUT/tc_check.log
NA_BENCH_LIGHT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python -c "import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['clocks'])"
