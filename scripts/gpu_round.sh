#!/bin/bash
# One GPU-box session: sanity, sanitizer, parity tests, bench, ncu.  Everything lands in gpurun_out/.
# usage: scripts/gpu_round.sh [tag] [steps...]   (steps default: all)
set -u
TAG=${1:-r1}; shift || true
STEPS=${*:-"info sanitize tests smoke bench launches ncu"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for s in $STEPS; do
  echo "=== $s ($(date +%T))"
  case $s in
    info) nvidia-smi > $OUT/nvidia_smi.txt 2>&1; nproc > $OUT/nproc.txt; python -c "import torch;print(torch.cuda.get_device_name(0))" ;;
    sanitize) timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $OUT/sanitize.log 2>&1; echo "sanitize exit $?"; tail -5 $OUT/sanitize.log ;;
    tests) timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/pytest_gpu.log ;;
    smoke) timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err ;;
    benchref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json ;;
    launches) NA_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "ncu launches exit $?"; tail -3 $OUT/launches.csv ;;
    ncu) NA_BENCH_LIGHT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 4 -c 2 -f -o $OUT/prof_mlp python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu full exit $?"; ls -la $OUT/ ;;
    tmsmall) NA_CHECK_MODES=tc timeout 150 python scripts/tc_check.py small > $OUT/tm_small.log 2>&1; echo "tm small exit $?"; tail -12 $OUT/tm_small.log ;;
    tmmixed) NA_CHECK_MODES=tc,tc_mixed timeout 150 python scripts/tc_check.py small > $OUT/tm_mixed.log 2>&1; echo "tm mixed exit $?"; tail -12 $OUT/tm_mixed.log ;;
    tmcheck) NA_CHECK_MODES=${MODES:-tc,tc_mixed,tc2acc} timeout 600 python scripts/tc_check.py > $OUT/tm_check.log 2>&1; echo "tm check exit $?"; tail -24 $OUT/tm_check.log ;;
    dram) NA_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mlp_t -c 70 --csv --log-file $OUT/dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/dram_bench.log 2>&1; echo "ncu dram exit $?"; tail -2 $OUT/dram.csv ;;
    tcsmall) timeout 180 python scripts/tc_check.py small > $OUT/tc_small.log 2>&1; echo "tc small exit $?"; tail -20 $OUT/tc_small.log ;;
    tcsan) timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/tc_check.py small > $OUT/tc_san.log 2>&1; echo "tc sanitize exit $?"; tail -30 $OUT/tc_san.log ;;
    tccheck) timeout 600 python scripts/tc_check.py > $OUT/tc_check.log 2>&1; echo "tc check exit $?"; tail -14 $OUT/tc_check.log ;;
    tccheck1) NA_TC_TWO_ACC=0 timeout 600 python scripts/tc_check.py > $OUT/tc_check_oneacc.log 2>&1; echo "tc check (one acc) exit $?"; tail -14 $OUT/tc_check_oneacc.log ;;
    testsfull) timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 $OUT/pytest_gpu.log ;;
    benchtc) NA_PRECISION=tc timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_tc.json 2> $OUT/bench_tc.err; echo "bench tc exit $?"; tail -c 3000 $OUT/bench_tc.json; tail -5 $OUT/bench_tc.err ;;
    teststc) NA_PRECISION=tc timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu_tc.log 2>&1; echo "pytest tc exit $?"; tail -40 $OUT/pytest_gpu_tc.log ;;
    ncutc) timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_t -c 2 -f -o $OUT/prof_mlp_tc python scripts/prof_mlp.py tc > $OUT/ncu_tc.log 2>&1; echo "ncu tc exit $?"; tail -3 $OUT/ncu_tc.log ;;
    launchestc) NA_PRECISION=tc NA_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_tc_bench.log 2>&1; echo "ncu launches exit $?"; tail -3 $OUT/launches_tc.csv ;;
    traintests) timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -x > $OUT/pytest_train.log 2>&1; echo "pytest train exit $?"; tail -40 $OUT/pytest_train.log ;;
    trainsan) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -m gpu -q -s -x -k "noeik or neus_backward" > $OUT/train_san.log 2>&1; echo "train sanitize exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" $OUT/train_san.log | head -20 ;;
    benchtrain) timeout 900 python bench.py --workload train --steps 2 --warmup 1 > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "bench train exit $?"; tail -c 2500 $OUT/bench_train.json; tail -5 $OUT/bench_train.err ;;
    launchestrain) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_train.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_train_bench.log 2>&1; echo "ncu launches train exit $?"; tail -2 $OUT/launches_train.csv ;;
    ncutrain) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd|wgrad" -s 2 -c 2 -f -o $OUT/prof_train python scripts/prof_train.py > $OUT/ncu_train.log 2>&1; echo "ncu train exit $?"; tail -3 $OUT/ncu_train.log ;;
    cliptests) timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q -s -x > $OUT/pytest_clip.log 2>&1; echo "pytest clip exit $?"; tail -30 $OUT/pytest_clip.log ;;
    clipsan) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_clip.py -m gpu -q -s -x -k "forward_and_image_gradient and 3" > $OUT/clip_san.log 2>&1; echo "clip sanitize exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" $OUT/clip_san.log | head -20 ;;
    traintestsfp32) NA_BWD=fp32 NA_WGRAD=fp32 timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s > $OUT/pytest_train_fp32.log 2>&1; echo "pytest train fp32 exit $?"; grep -E "worst|passed|failed|ln_beta" $OUT/pytest_train_fp32.log | tail -12 ;;
    *) echo "unknown step $s" ;;
  esac
done
echo "=== done ($(date +%T))"
