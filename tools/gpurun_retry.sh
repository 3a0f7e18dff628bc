#!/bin/bash
# gpurun with retries while the pod answers busy/transient (exit 3 / "transient").   tools/gpurun_retry.sh <timeout> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|no box\|busy"; then sleep 60; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
