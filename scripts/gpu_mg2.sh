#!/bin/bash
# 2-GPU session: render bench (config 2), train bench (NCCL all-reduce of the packed gradient, rank-synchronised inputs)
OUT=gpurun_out/mg2; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== render x2 ($(date +%T))"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_render_x2.json 2> $OUT/bench_render_x2.err; echo "rc=$?"; tail -1 $OUT/bench_render_x2.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'])"
echo "=== train x2 ($(date +%T))"
timeout 900 $TR --master-port 29542 bench.py --workload train --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train_x2.json 2> $OUT/bench_train_x2.err; echo "rc=$?"; tail -1 $OUT/bench_train_x2.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['phases_ms'])"; tail -3 $OUT/bench_train_x2.err
echo "=== train x2, interleaved tiles ($(date +%T))"
NA_PARTITION=tiles timeout 900 $TR --master-port 29543 bench.py --workload train --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_train_x2_tiles.json 2> $OUT/bench_train_x2_tiles.err; echo "rc=$?"; tail -1 $OUT/bench_train_x2_tiles.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['phases_ms'])"
echo "=== done ($(date +%T))"
