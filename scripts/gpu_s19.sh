#!/bin/bash
# s19: validation of the folded-scale SO/TR epilogues + cycle counters out of the default build: full GPU suite, kernel probes, benches
OUT=gpurun_out/s19; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== tc_check ($(date +%T))"
NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check.log 2>&1; grep -E "^tc|CTA0|Linf" $OUT/tc_check.log
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 4 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'], d['style_ms_per_step'])"
echo "=== render bench ($(date +%T))"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python -c "import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak'], d['train_probe']['volsdf'])"
echo "=== done ($(date +%T))"
