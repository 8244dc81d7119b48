#!/bin/bash
# 2-GPU check of the style-loss phase: per-step event time vs host wall time
OUT=gpurun_out/mg2b; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nproc; free -g | head -2
for part in block tiles; do
echo "=== train x2 $part ($(date +%T))"
NA_PARTITION=$part timeout 900 $TR --master-port $((29000 + RANDOM % 900)) bench.py --workload train --gpus 2 --steps 5 --warmup 1 --no-cpu-baseline > $OUT/bench_train_x2_$part.json 2> $OUT/bench_train_x2_$part.err; echo "rc=$?"; tail -1 $OUT/bench_train_x2_$part.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['style_ms_per_step'], d['style_host_ms_per_step'])"
done
echo "=== done ($(date +%T))"
