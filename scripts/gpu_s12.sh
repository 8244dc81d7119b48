#!/bin/bash
OUT=gpurun_out/s12; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for v in "" _noguard ""; do
  echo "=== variant '${v}' ($(date +%T))"
  NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so NA_CHECK_MODES=tc timeout 300 python scripts/tc_check.py > $OUT/tc_check$v.log 2>&1; grep -E "^tc |CTA0" $OUT/tc_check$v.log
done
echo "=== render bench ($(date +%T))"
for v in "" _noguard; do
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench$v.json 2> $OUT/bench$v.err; python -c "import json; d=json.load(open('$OUT/bench$v.json')); print('$v', d['ms_per_step'], d['value'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak'])"
done
echo "=== done ($(date +%T))"
