#!/bin/bash
OUT=gpurun_out/s11; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== tc_check ($(date +%T))"
NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check.log 2>&1; grep -E "^tc|CTA0" $OUT/tc_check.log
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "=== stall matrix ($(date +%T))"
scripts/hang_matrix.sh $OUT/hang 5 "lazy_nopre_d1:CUDA_MODULE_LOADING=LAZY,NA_PRELOAD=0,NA_PROBE_DIAG=1" | cut -c1-140
echo "=== render bench ($(date +%T))"
for prec in tc_mixed tc; do
NA_PRECISION=$prec timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$prec.json 2> $OUT/bench_$prec.err; python -c "import json; d=json.load(open('$OUT/bench_$prec.json')); print('$prec', d['ms_per_step'], d['value'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak'])"
done
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"
echo "=== done ($(date +%T))"
