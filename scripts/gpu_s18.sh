#!/bin/bash
# s18: SO-pass stamps; default bench line (with the small-beta probe); per-step style-loss times
OUT=gpurun_out/s18; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== bw trace ($(date +%T))"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200_trace.so timeout 300 python scripts/bw_trace.py $OUT/bw_trace.npy 2>&1 | tail -1
python scripts/trace_show.py $OUT/bw_trace.npy 41 | tail -1
python - <<PY
import numpy as np
d=np.load('$OUT/bw_trace.npy').astype(np.int64)[16:]
E=d[352:352+44*12].reshape(44,4,3); x=d[880:896]
print('SO g=28 pass 1: start->Dready', E[28,1,0]-x[7], 'ld', x[10]-E[28,1,0], 'decode+math', x[11]-x[10], 'stash16', x[12]-x[11], 'qstore16', x[13]-x[12], 'store_a16', x[14]-x[13])
PY
echo "=== default bench ($(date +%T))"
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']); print(d['small_beta_probe']); print(d['cpu_baseline']); print(d['reference_gpu']); print(d['train_probe'])"
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 4 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'], d['style_ms_per_step'])"
echo "=== done ($(date +%T))"
