#!/bin/bash
OUT=gpurun_out/s5; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== wgrad unit test ($(date +%T))"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k wgrad_f16 > $OUT/pytest_wgrad.log 2>&1; echo "rc=$?"; grep -E "wgrad_f16 m=|passed|failed|Error|error" $OUT/pytest_wgrad.log | head -20
echo "=== train tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "not wgrad_f16" > $OUT/pytest_train.log 2>&1; echo "rc=$?"; grep -E "worst|Trainer.forward|passed|failed|FAILED|Error" $OUT/pytest_train.log | cut -c1-300 | tail -24
echo "=== full suite ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_gpu.log | cut -c1-300
echo "=== bench train ($(date +%T))"
timeout 600 python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"; tail -3 $OUT/bench_train.err
echo "=== launches of two patches ($(date +%T))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/launches_train_patch.csv python scripts/prof_train.py > $OUT/prof_train.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/s5/launches_train_patch.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r)); k=d['Kernel Name'][:70]; m=d['Metric Name']; v=float(d['Metric Value'].replace(',',''))
    a=agg.setdefault(k,collections.defaultdict(float)); a[m]+=v; a['n_'+m]+=1
for k,a in agg.items():
    n=a['n_gpu__time_duration.sum']
    print(f"{k:70s} n={int(n):3d} t={a['gpu__time_duration.sum']/1e6:8.3f} ms  rd={a['dram__bytes_read.sum']/1e9:7.3f} GB wr={a['dram__bytes_write.sum']/1e9:7.3f} GB tensor%={a['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']/max(n,1):5.1f}")
PY
echo "=== render bench default + tc_mixed ($(date +%T))"
for prec in tc tc_mixed; do
NA_PRECISION=$prec timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$prec.json 2> $OUT/bench_$prec.err; python -c "import json; d=json.load(open('$OUT/bench_$prec.json')); print('$prec', d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
done
echo "=== done ($(date +%T))"
