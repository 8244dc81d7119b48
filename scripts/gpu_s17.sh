#!/bin/bash
# s16: merged epilogue kinds in the training program (code size 17.5 k -> 14.7 k SASS instructions); extra stamps around write_vbar0
OUT=gpurun_out/s17; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== bw trace ($(date +%T))"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200_trace.so timeout 300 python scripts/bw_trace.py $OUT/bw_trace.npy 2>&1 | tail -2
python scripts/trace_show.py $OUT/bw_trace.npy 41 | tail -4
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_train.log | cut -c1-300
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
echo "=== emulated 8-GPU share on one GPU ($(date +%T))"
for emu in; do
NA_BENCH_LIGHT=1 NA_BENCH_EMULATE_WORLD=$emu timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_emu$emu.json 2> $OUT/bench_emu$emu.err; python -c "import json; d=json.load(open('$OUT/bench_emu$emu.json')); print('emu $emu', d['ms_per_step'], d['clocks'])"
done
echo "=== done ($(date +%T))"
