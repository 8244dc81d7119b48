#!/bin/bash
# s31: validation of the tree: full GPU suite, smoke, default bench (both arms), train bench, sanitizers on the split programs
OUT=gpurun_out/s31; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "=== smoke ($(date +%T))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/smoke.log | cut -c1-300
echo "=== default bench ($(date +%T))"
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak']); print(d['small_beta_probe']); print(d['cpu_baseline']); print(d['reference_gpu']); print(d['train_probe'])"
echo "=== reference arm ($(date +%T))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cut -c1-400 $OUT/bench_reference.json
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
echo "=== sanitizers on the split programs ($(date +%T))"
for tool in memcheck synccheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "split and (tc_mixed-False or neus and tc_mixed)" > $OUT/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
echo "=== done ($(date +%T))"
