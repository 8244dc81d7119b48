#!/bin/bash
OUT=gpurun_out/s7; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_dropin_gpu.py -m gpu -q -s > $OUT/pytest_train.log 2>&1; echo "rc=$?"; grep -E "worst|Trainer.forward|passed|failed|FAILED|Error" $OUT/pytest_train.log | cut -c1-250 | tail -24
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"; tail -3 $OUT/bench_train.err
echo "=== launches of two patches ($(date +%T))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/launches_train_patch.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/s7/launches_train_patch.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); k=d['Kernel Name'][:60]; m=d['Metric Name']; v=float(d['Metric Value'].replace(',',''))
    a=agg.setdefault(k,collections.defaultdict(float)); a[m]+=v; a['n_'+m]+=1
for k,a in agg.items():
    n=a['n_gpu__time_duration.sum']
    if a['gpu__time_duration.sum'] < 30000: continue
    print(f"{k:60s} n={int(n):3d} t={a['gpu__time_duration.sum']/1e6:8.3f} ms  rd={a['dram__bytes_read.sum']/1e9:7.3f} GB wr={a['dram__bytes_write.sum']/1e9:7.3f} GB tensor%={a['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']/max(n,1):5.1f}")
PY
echo "=== ncu full: BW kernel ($(date +%T))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tmem -s 8 -c 1 -f -o $OUT/prof_bw python scripts/prof_train.py > $OUT/ncu_bw.log 2>&1; echo "ncu full rc=$?"; ls -la $OUT/*.ncu-rep
echo "=== done ($(date +%T))"
