#!/bin/bash
OUT=gpurun_out/s9; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== full suite, default tc_mixed ($(date +%T))"
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" $OUT/pytest_gpu.log | tail -8 | cut -c1-300
echo "=== smoke ($(date +%T))"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== ncu full: render kernels tc_mixed and tc ($(date +%T))"
for prec in tc_mixed tc; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tmem -c 2 -f -o $OUT/prof_mlp_$prec python scripts/prof_mlp.py $prec > $OUT/ncu_mlp_$prec.log 2>&1; echo "ncu $prec rc=$?"
done
echo "=== launch list of the bench ($(date +%T))"
NA_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "ncu launches rc=$?"
echo "=== render bench default ($(date +%T))"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frame_frac_of_peak'], d['cpu_baseline'], d['reference_gpu'])"; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; python -c "
import json; d=json.load(open('$OUT/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline'])"
echo "=== bench train default ($(date +%T))"
timeout 600 python bench.py --workload train --steps 2 --warmup 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_train.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['cpu_baseline'])"; tail -2 $OUT/bench_train.err
echo "=== done ($(date +%T))"
