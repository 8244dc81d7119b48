#!/bin/bash
# 2-GPU session c: final tree -- render, VolSDF and NeuS fine-tune steps over NCCL (split training programs, launch groups over non-contiguous patches)
OUT=gpurun_out/mg2c; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== render x2 ($(date +%T))"
timeout 600 $TR --master-port 29641 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_render_x2.json 2> $OUT/bench_render_x2.err; echo "rc=$?"; tail -1 $OUT/bench_render_x2.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'])"
for fw in volsdf neus; do
echo "=== train x2 $fw ($(date +%T))"
timeout 900 $TR --master-port 2964$((RANDOM % 5 + 2)) bench.py --workload train --framework $fw --gpus 2 --steps 4 --warmup 2 --no-cpu-baseline > $OUT/bench_train_${fw}_x2.json 2> $OUT/bench_train_${fw}_x2.err; echo "rc=$?"; tail -1 $OUT/bench_train_${fw}_x2.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], {k[:12]: round(v,1) for k,v in d['phases_ms'].items()}, d['style_ms_per_step'])"; tail -2 $OUT/bench_train_${fw}_x2.err | cut -c1-200
done
echo "=== done ($(date +%T))"
