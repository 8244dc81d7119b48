#!/bin/bash
# 4-GPU session: config 2 render and the VolSDF fine-tune step on the final tree
OUT=gpurun_out/mg4; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29741 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_render_x4.json 2> $OUT/bench_render_x4.err; echo "rc=$?"; tail -1 $OUT/bench_render_x4.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'])"
timeout 300 $TR --master-port 29742 bench.py --workload train --gpus 4 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_train_x4.json 2> $OUT/bench_train_x4.err; echo "rc=$?"; tail -1 $OUT/bench_train_x4.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()})"
