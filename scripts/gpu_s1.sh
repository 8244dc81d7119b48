#!/bin/bash
# session 1: instrumented build sanity, stall matrix, sanitizers
OUT=gpurun_out/s1; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi > $OUT/nvidia_smi.txt 2>&1; nproc > $OUT/nproc.txt
echo "=== pytest ($(date +%T))"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "=== matrix ($(date +%T))"
scripts/hang_matrix.sh $OUT/hang 8 "lazy_nopre_d2:CUDA_MODULE_LOADING=LAZY,NA_PRELOAD=0,NA_PROBE_DIAG=2"
scripts/hang_matrix.sh $OUT/hang 6 "lazy_nopre_d1:CUDA_MODULE_LOADING=LAZY,NA_PRELOAD=0,NA_PROBE_DIAG=1"
scripts/hang_matrix.sh $OUT/hang 6 "lazy_pre_d1:CUDA_MODULE_LOADING=LAZY,NA_PRELOAD=1,NA_PROBE_DIAG=1"
echo "=== sanitizers ($(date +%T))"
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > $OUT/san_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:" $OUT/san_$tool.log | tail -5
done
echo "=== done ($(date +%T))"
