#!/bin/bash
# s36: compute-sanitizer over the render parity tests and the NeuS / VolSDF Trainer.forward tests on the final build
OUT=gpurun_out/s36; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for tool in memcheck synccheck; do
echo "=== $tool: render parity ($(date +%T))"
timeout 420 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "networks or ragged or baseline_config or neus_render_vs or white_background" > $OUT/sanitizer_${tool}_render.log 2>&1; echo "rc=$?"; grep -aE "ERROR SUMMARY|passed|failed" $OUT/sanitizer_${tool}_render.log | tail -2
done
echo "=== memcheck: Trainer.forward (split programs, launch groups) ($(date +%T))"
timeout 420 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "trainer_forward and tc_mixed" > $OUT/sanitizer_memcheck_trainer.log 2>&1; echo "rc=$?"; grep -aE "ERROR SUMMARY|passed|failed" $OUT/sanitizer_memcheck_trainer.log | tail -2
echo "=== done ($(date +%T))"
