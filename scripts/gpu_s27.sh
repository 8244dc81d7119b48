#!/bin/bash
# s28 (has_rad-specialised backward-only kernels, slim epilogue context): dedicated backward-only instantiation (B2) + four-columns-at-a-time SO / TR arithmetic
OUT=gpurun_out/s27; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
echo "=== ncu ($(date +%T))"
timeout 600 ncu --metrics $M --clock-control none -k mlp_tmem_kernel --csv --log-file $OUT/ncu.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/ncu.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in d.items():
    if m['gpu__time_duration.sum'] > 1.2e6: print(i, k, {a.split('.')[0][-30:]: round(b, 2) for a, b in m.items()})
PY
echo "=== split tests ($(date +%T))"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_train.log | cut -c1-200
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
echo "=== done ($(date +%T))"
