#!/bin/bash
# s33: hoisted run-time tests in the merged epilogue kinds; final ncu captures of the training kernels and the render kernels
OUT=gpurun_out/s33; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_dropin_gpu.py -m gpu -q > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_train.log | cut -c1-200
echo "=== launch list ($(date +%T))"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_train_patch.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/launches_train_patch.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:44]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in list(d.items())[-40:]:
    if m['gpu__time_duration.sum'] > 5e5: print(i, k, {a.split('.')[0][-26:]: round(b, 1) for a, b in m.items()})
PY
echo "=== full captures ($(date +%T))"
timeout 600 ncu --set full --clock-control none --import-source on -k mlp_tmem_kernel -s 16 -c 2 -o $OUT/prof_split_halves python scripts/prof_train.py > $OUT/ncu_split_halves.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k mlp_tmem_kernel -c 2 -o $OUT/prof_mlp_tc_mixed python scripts/prof_mlp.py tc_mixed > $OUT/ncu_mlp.log 2>&1; echo "rc=$?"
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print('train', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
ls -la $OUT | head -20
echo "=== done ($(date +%T))"
