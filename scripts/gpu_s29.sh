#!/bin/bash
# s29: g-plane rows one pass ahead in the specialised backward-only kernel (_pg) vs default
OUT=gpurun_out/s29; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for v in "" _pg; do
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 ncu --metrics $M --clock-control none -k mlp_tmem_kernel --csv --log-file $OUT/ncu$v.csv python scripts/prof_train.py > $OUT/prof_train$v.log 2>&1; echo "variant '$v' rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/ncu$v.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in d.items():
    if m['gpu__time_duration.sum'] > 1.8e6: print(i, k, {a.split('.')[0][-30:]: round(b, 2) for a, b in m.items()})
PY
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -k "split" 2>&1 | tail -1
done
