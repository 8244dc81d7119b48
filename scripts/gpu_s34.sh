#!/bin/bash
# s34: NeuS fine-tune step (BASELINE configs[2]) through Trainer.forward, split vs one-launch programs
OUT=gpurun_out/s34; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for sp in 1 0; do
NA_BW_SPLIT=$sp timeout 600 python bench.py --workload train --framework neus --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train_neus_split$sp.json 2> $OUT/bench_train_neus_split$sp.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train_neus_split$sp.json')); print('neus split $sp', d['ms_per_step'], d['value'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"; tail -2 $OUT/bench_train_neus_split$sp.err
done
