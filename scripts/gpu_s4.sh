#!/bin/bash
OUT=gpurun_out/s4; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== dropin train test ($(date +%T))"
timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x > $OUT/pytest_dropin.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_dropin.log | cut -c1-400
echo "=== path divergence / sdf error per mode ($(date +%T))"
for cfg in "tc:2:0" "tc:2:5" "tc:2:6" "tc:2:7" "tc:2:8" "tc:1:4.3" "tc:0:14.3" "tc_mixed:2:6"; do
  IFS=: read prec order deb <<< "$cfg"
  NA_TM_ORDER=${order} NA_TM_DEBIAS=${deb} timeout 300 python scripts/path_div.py $prec 2>/dev/null | tail -1 | tee -a $OUT/path_div.jsonl
done
echo "=== both nets, default order 2 ($(date +%T))"
for deb in 0 6; do NA_TM_DEBIAS=$deb NA_CHECK_MODES=tc,tc2acc timeout 300 python scripts/tc_check.py small 2>&1 | grep Linf | sed "s/^/debias=$deb /"; done | tee $OUT/tc_small.log
echo "=== speed order 2 ($(date +%T))"
NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check_order2.log 2>&1; grep -E "^tc|CTA0" $OUT/tc_check_order2.log
echo "=== bench ($(date +%T))"
for prec in tc tc_mixed; do
NA_PRECISION=$prec timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$prec.json 2> $OUT/bench_$prec.err; python -c "import json; d=json.load(open('$OUT/bench_$prec.json')); print('$prec', d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
done
echo "=== racecheck ($(date +%T))"
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > $OUT/san_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|smoke:" $OUT/san_racecheck.log | tail -4
echo "=== done ($(date +%T))"
