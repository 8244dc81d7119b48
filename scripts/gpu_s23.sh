#!/bin/bash
# s23: softplus' codes one pass ahead in registers (default: also across GEMMs; _nos0: within a GEMM only)
OUT=gpurun_out/s23; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for v in "" _nos0; do
echo "=== variant '$v' ($(date +%T))"
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 ncu --metrics $M --clock-control none -k mlp_tmem_kernel --csv --log-file $OUT/ncu$v.csv python scripts/prof_train.py > $OUT/prof_train$v.log 2>&1; echo "rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$OUT/ncu$v.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]: d.setdefault((int(r[ii]), r[ki][:34]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in d.items():
    if m['gpu__time_duration.sum'] > 1.4e6: print(i, k, {a.split('.')[0][-30:]: round(b, 2) for a, b in m.items()})
PY
done
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_train.log | cut -c1-300
echo "=== bench train ($(date +%T))"
for v in "" _nos0; do
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 python bench.py --workload train --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train$v.json 2> $OUT/bench_train$v.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train$v.json')); print('lib$v', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
echo "=== done ($(date +%T))"
