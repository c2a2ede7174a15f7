#!/bin/bash
# A/B of the current library against the build of the previous commit (tools/bin/libsoftspoken_b200_prev.so, made by
# checking HEAD out into a scratch worktree): whole-classifier time over 1,005 windows in alternating processes on one
# box, logits compared bit for bit; conv1_direct through the launch profile of both.  Output: gpurun_out/ab_prev.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_prev.txt
: > $OUT
PREV=$PWD/tools/bin/libsoftspoken_b200_prev.so
rm -f /tmp/ab_ref_*.pt
for mode in f16x3 bf16; do
  for round in 1 2; do
    SOFTSPOKEN_B200_LIB=$PREV python tools/time_classify.py $mode 1005 /tmp/ab_ref_$mode.pt 2>&1 | sed "s/^/prev /" >> $OUT
    python tools/time_classify.py $mode 1005 /tmp/ab_ref_$mode.pt 2>&1 | sed "s/^/new  /" >> $OUT
  done
done
cat $OUT
