#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in ${CFGS:-"1 1 1" "1 1 0" "1 1 5" "1 1 3" "1 2 1" "0 1 1"}; do
  set -- $cfg
  AVSEP_DFT_FUSED_STAGE=$1 AVSEP_DFT_FOLD3_GROUPS=$2 AVSEP_DFT_FOLD_PREFETCH=$3 timeout 90 python tools/time_stft.py 2>&1 | tail -1
done
