#!/bin/bash
# s14: L2 prefetch of the next GEMM's scratch (default build) vs -DNA_TM_NO_PREFETCH (_nopf); patch launch groups (NA_PATCH_GROUP)
OUT=gpurun_out/s14; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for v in "" _nopf; do
  echo "=== variant '${v}' tc_check ($(date +%T))"
  NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so NA_CHECK_MODES=tc,tc_mixed timeout 300 python scripts/tc_check.py > $OUT/tc_check$v.log 2>&1; grep -E "^tc|CTA0" $OUT/tc_check$v.log
  echo "=== variant '${v}' ncu BW program ($(date +%T))"
  NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"mlp_tmem|wgrad" --csv --log-file $OUT/ncu_bw$v.csv python scripts/prof_train.py > $OUT/prof_train$v.log 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open('$OUT/ncu_bw$v.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
d = {}
for r in rows[1:]: d.setdefault((r[ii], r[ki][:40]), {})[r[mi]] = r[vi]
for k, m in d.items():
    if float(m['gpu__time_duration.sum'].replace(',', '')) > 3e5: print(k, {a.split('.')[0][-28:]: b for a, b in m.items()})
PY
done
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_surface_render.py -m gpu -q -x > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_train.log | cut -c1-300
echo "=== bench train ($(date +%T))"
for v in "" _nopf; do for grp in 6 1; do
NA_PATCH_GROUP=$grp NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 600 python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train${v}_g$grp.json 2> $OUT/bench_train${v}_g$grp.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train${v}_g$grp.json')); print('lib$v group $grp', d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])"
done; done
echo "=== render bench ($(date +%T))"
for v in "" _nopf; do
NA_LIB_PATH=$PWD/nerf-art_b200/libnerfart_b200$v.so timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench$v.json 2> $OUT/bench$v.err; python -c "import json; d=json.load(open('$OUT/bench$v.json')); print('$v', d['ms_per_step'], d['value'], d['clocks'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak'])"
done
echo "=== done ($(date +%T))"
