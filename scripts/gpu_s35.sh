#!/bin/bash
# s35: last validation of the round's tree: full GPU suite, smoke, default bench (driver contract), 2-rank check is scripts/gpu_mg2.sh
OUT=gpurun_out/s35; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "=== smoke ($(date +%T))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -4 $OUT/smoke.log | cut -c1-300
echo "=== default bench ($(date +%T))"
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak']); print(d['train_probe'])"
echo "=== done ($(date +%T))"
