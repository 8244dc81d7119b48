#!/bin/bash
# s22b: ncu --set full of the two halves of the split training program (second patch: launches 16 = forward with stash, 17 = backward only)
OUT=gpurun_out/s22; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k mlp_tmem_kernel -s 16 -c 2 -o $OUT/prof_split_halves python scripts/prof_train.py > $OUT/ncu_split_halves.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_split_halves.log
ls -la $OUT
