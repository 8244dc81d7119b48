#!/bin/bash
OUT=gpurun_out/s8; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log | cut -c1-300
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"; tail -3 $OUT/bench_train.err
echo "=== launches of two patches ($(date +%T))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file $OUT/launches_train_patch.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/s8/launches_train_patch.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); k=d['Kernel Name'][:60]; m=d['Metric Name']; v=float(d['Metric Value'].replace(',',''))
    a=agg.setdefault(k,collections.defaultdict(float)); a[m]+=v; a['n_'+m]+=1
for k,a in agg.items():
    n=a['n_gpu__time_duration.sum']
    if a['gpu__time_duration.sum'] < 30000: continue
    print(f"{k:60s} n={int(n):3d} t={a['gpu__time_duration.sum']/1e6:8.3f} ms  rd={a['dram__bytes_read.sum']/1e9:7.3f} GB wr={a['dram__bytes_write.sum']/1e9:7.3f} GB tensor%={a['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']/max(n,1):5.1f} l2hit={a['lts__t_sector_hit_rate.pct']/max(n,1):5.1f}")
PY
echo "=== tc_mixed evaluation ($(date +%T))"
timeout 600 python scripts/mixed_eval.py 2>/dev/null | tee $OUT/mixed_eval.jsonl | cut -c1-260
echo "=== render bench (full line) ($(date +%T))"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline'], d['reference_gpu'], d['train_probe'])"; tail -3 $OUT/bench.err
echo "=== done ($(date +%T))"
