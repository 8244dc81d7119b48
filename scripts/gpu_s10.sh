#!/bin/bash
OUT=gpurun_out/s10; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for v in "" _noguard _nobar _order0 _all3 ""; do
  echo "=== variant '${v}' ($(date +%T))"
  NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so NA_CHECK_MODES=tc timeout 300 python scripts/tc_check.py > $OUT/tc_check$v.log 2>&1; grep -E "^tc |CTA0" $OUT/tc_check$v.log
done
echo "=== done ($(date +%T))"
