#!/bin/bash
# s32: 16-bit fixed-point softplus' codes (default) vs the 15-bit exp codes (_c15): accuracy, kernel times, tests, benches
OUT=gpurun_out/s32; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in "" _c15; do
echo "=== variant '$v' ($(date +%T))"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check$v.log 2>&1; grep -E "^tc|Linf" $OUT/tc_check$v.log | grep -v fp32
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 ncu --metrics $M --clock-control none -k mlp_tmem_kernel --csv --log-file $OUT/ncu$v.csv python scripts/prof_train.py > $OUT/prof_train$v.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/ncu$v.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in list(d.items())[9:]:
    if m['gpu__time_duration.sum'] > 1.2e6: print(i, k, {a.split('.')[0][-30:]: round(b, 2) for a, b in m.items()})
PY
done
echo "=== full suite, default ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
echo "=== benches ($(date +%T))"
for v in "" _c15; do
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so NA_BENCH_LIGHT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench$v.json 2> $OUT/bench$v.err; python -c "import json; d=json.load(open('$OUT/bench$v.json')); print('render$v', d['ms_per_step'], d['clocks']['sm_mhz'])"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train$v.json 2> $OUT/bench_train$v.err; python -c "
import json; d=json.load(open('$OUT/bench_train$v.json')); print('train$v', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
echo "=== done ($(date +%T))"
