#!/bin/bash
# Split-K sub-accumulation: accuracy (logits vs float64 truth) and classifier time per policy (SS_TC_SUB = groups in the
# layers at >= 64 x 128, SS_TC_SUB_DEEP = groups in the deeper layers).
for cfg in "1 1" "1 2" "1 4" "1 8" "1 16" "2 16"; do
  set -- $cfg
  echo "== SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2"
  SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2 python tools/time_classify.py f16x3 1005
  SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2 python tools/precision_study.py | grep "f16x3"
done
