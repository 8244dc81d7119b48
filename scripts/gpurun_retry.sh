#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command...>  -- retries while the pod answers "busy" (exit 3 / status=transient)
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then echo "[retry $attempt] pod busy, waiting"; sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
