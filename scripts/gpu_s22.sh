#!/bin/bash
# s22: ncu evidence for the split training program (launch list + one full capture of each half) and the render kernel
OUT=gpurun_out/s22; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
echo "=== launch list, two patches ($(date +%T))"
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_train_patch.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/launches_train_patch.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:48]), {})[r[mi]] = float(r[vi].replace(',', ''))
tot = collections.Counter(); cnt = collections.Counter()
for (i, k), m in d.items(): tot[k] += m['gpu__time_duration.sum']; cnt[k] += 1
for k, v in tot.most_common(12): print(f'{v/1e6:9.3f} ms {cnt[k]:4d}x  {k}')
for (i, k), m in d.items():
    if m['gpu__time_duration.sum'] > 5e5: print(i, k, {a.split('.')[0][-26:]: round(b, 1) for a, b in m.items()})
PY
echo "=== full captures ($(date +%T))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"1, \(bool\)1, \(bool\)1" -s 1 -c 1 -o $OUT/prof_bw_only python scripts/prof_train.py > $OUT/ncu_bw_only.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"1, \(bool\)1, \(bool\)0" -s 1 -c 1 -o $OUT/prof_fwd_stash python scripts/prof_train.py > $OUT/ncu_fwd_stash.log 2>&1; echo "rc=$?"
ls -la $OUT
echo "=== done ($(date +%T))"
