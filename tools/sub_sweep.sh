#!/bin/bash
# Split-K sub-accumulation: accuracy (logits vs float64 truth) and classifier time per policy (SS_TC_SUB = groups in the
# layers at >= 64 x 128, SS_TC_SUB_DEEP = groups in the deeper layers); "nosub" = library built with -DSS_TC_SUBACC=0.
echo "== nosub build"; SOFTSPOKEN_B200_LIB=tools/bin/libss_nosub.so python tools/time_classify.py f16x3 1005
for cfg in "1 1" "1 16" "2 16" "3 16" "4 16" "8 16" "16 16"; do
  set -- $cfg
  echo "== SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2"
  SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2 python tools/time_classify.py f16x3 1005
  SS_TC_SUB=$1 SS_TC_SUB_DEEP=$2 python tools/precision_study.py | grep "f16x3"
done
