#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3, nothing charged).
# usage: [GPUS=N] gpurun_retry.sh <timeout> '<command>'
t=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $extra --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
