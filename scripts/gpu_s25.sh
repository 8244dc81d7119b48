#!/bin/bash
# s25: NeuS split training program; all training tests; NeuS / VolSDF train probes of the main bench line
OUT=gpurun_out/s25; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "=== split tests ($(date +%T))"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "split" -s > $OUT/pytest_split.log 2>&1; echo "rc=$?"; grep -E "split vs|passed|failed|Error|error" $OUT/pytest_split.log | head -20
echo "=== train + dropin tests ($(date +%T))"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_dropin_gpu.py -m gpu -q > $OUT/pytest_train.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_train.log | cut -c1-300
echo "=== train probes ($(date +%T))"
timeout 600 python - > $OUT/probe.log 2>&1 <<PY
import json, os, torch, bench
for sp in ('1', '0'):
    os.environ['NA_BW_SPLIT'] = sp
    import nerfart_b200
    if sp == '0':
        # one-launch programs: render_patch without the stash
        src = open('bench.py').read().replace("train_stash=precision in ('tc', 'tc_mixed')", "train_stash=False")
        ns = {}; exec(compile(src, 'bench_nosplit', 'exec'), ns); probe = ns['train_probe']
    else:
        probe = bench.train_probe
    r = probe(torch.device('cuda:0'), 'tc_mixed')
    print('split', sp, {k: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != 'backward_roofline'} for k, v in r.items()})
PY
cat $OUT/probe.log | tail -4
echo "=== done ($(date +%T))"
