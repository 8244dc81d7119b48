#!/bin/bash
# 8-GPU session b: fine-tune step (BASELINE config 5 shape) with the split training program, two warm-up steps
OUT=gpurun_out/mg8b; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nproc
echo "=== train x8 ($(date +%T))"
timeout 900 $TR --master-port 29563 bench.py --workload train --gpus 8 --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train_x8.json 2> $OUT/bench_train_x8.err; echo "rc=$?"; tail -1 $OUT/bench_train_x8.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['style_ms_per_step'], d['clocks'])"; tail -2 $OUT/bench_train_x8.err
echo "=== done ($(date +%T))"
