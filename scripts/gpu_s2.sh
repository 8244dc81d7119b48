#!/bin/bash
# session 2: after the mbarrier phase-aliasing fix -- stall matrix, 20 fresh train benches, sanitizers, tests
OUT=gpurun_out/s2; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== pytest ($(date +%T))"
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "Trainer.forward|passed|failed" $OUT/pytest_gpu.log | tail -8
echo "=== matrix ($(date +%T))"
scripts/hang_matrix.sh $OUT/hang 14 "lazy_nopre_d1:CUDA_MODULE_LOADING=LAZY,NA_PRELOAD=0,NA_PROBE_DIAG=1" | cut -c1-160
scripts/hang_matrix.sh $OUT/hang 6 "default_d0:NA_PROBE_DIAG=0" | cut -c1-160
echo "=== 20 fresh train benches, CUDA_MODULE_LOADING unset ($(date +%T))"
ok=0
for i in $(seq 1 20); do
  NA_BENCH_WATCHDOG_S=120 timeout 200 python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bt_$i.json 2> $OUT/bt_$i.err
  rc=$?; if [ $rc -eq 0 ] && grep -q '"value"' $OUT/bt_$i.json; then ok=$((ok+1)); fi
  echo "train bench $i rc=$rc $(python -c "import json,sys; d=json.load(open('$OUT/bt_$i.json')); print(round(d['ms_per_step']), 'ms')" 2>/dev/null)"
done
echo "train benches completed: $ok / 20"
echo "=== sanitizers ($(date +%T))"
for tool in synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > $OUT/san_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:" $OUT/san_$tool.log | tail -5
done
echo "=== done ($(date +%T))"
