#!/bin/bash
# 8-GPU session: BASELINE configs[3] (960x540x256 render over 8 GPUs) and configs[4] (fine-tune step over 8 GPUs); configs[1] at 8 GPUs is the
# driver's own scaling run
OUT=gpurun_out/mg8; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "=== config 4 render x8 ($(date +%T))"
timeout 600 $TR --master-port 29551 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_config4_x8.json 2> $OUT/bench_config4_x8.err; echo "rc=$?"; tail -1 $OUT/bench_config4_x8.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['config']['workload'][:80])"; tail -2 $OUT/bench_config4_x8.err
echo "=== train x8 (config 5) ($(date +%T))"
timeout 900 $TR --master-port 29553 bench.py --workload train --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train_x8.json 2> $OUT/bench_train_x8.err; echo "rc=$?"; tail -1 $OUT/bench_train_x8.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['phases_ms'])"; tail -3 $OUT/bench_train_x8.err
echo "=== done ($(date +%T))"
