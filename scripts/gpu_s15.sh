#!/bin/bash
# s15: per-GEMM clock trace of the training program; BW forward in tc_mixed; style-loss host profile; emulated 8-GPU share
OUT=gpurun_out/s15; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== bw trace ($(date +%T))"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200_trace.so timeout 300 python scripts/bw_trace.py $OUT/bw_trace.npy 2>&1 | tail -2
NA_BW_FWD=tc NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200_trace.so timeout 300 python scripts/bw_trace.py $OUT/bw_trace_fwdtc.npy 2>&1 | tail -1
echo "=== train + clip tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_clip.py -m gpu -q -x > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_train.log | cut -c1-300
grep -h "worst relative" $OUT/pytest_train.log | head
echo "=== style profile ($(date +%T))"
timeout 300 python scripts/prof_style.py > $OUT/prof_style.log 2>&1; head -12 $OUT/prof_style.log
echo "=== bench train ($(date +%T))"
for fw in "" tc; do
NA_BW_FWD=$fw timeout 600 python bench.py --workload train --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_train_$fw.json 2> $OUT/bench_train_$fw.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train_$fw.json')); print('bw fwd [$fw]', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
echo "=== emulated 8-GPU share on one GPU ($(date +%T))"
for emu in 1 8; do
NA_BENCH_LIGHT=1 NA_BENCH_EMULATE_WORLD=$emu timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_emu$emu.json 2> $OUT/bench_emu$emu.err; python -c "import json; d=json.load(open('$OUT/bench_emu$emu.json')); print('emu $emu', d['ms_per_step'], d['clocks'])"
done
echo "=== done ($(date +%T))"
