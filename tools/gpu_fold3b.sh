#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "1 1 1" "1 2 1" "1 1 3" "1 2 3" "0 1 1"; do
  set -- $cfg
  AVSEP_DFT_FUSED_STAGE=$1 AVSEP_DFT_FOLD3_SETS=$2 AVSEP_DFT_FOLD_PREFETCH=$3 timeout 90 python tools/time_stft.py 2>&1 | tail -1
done
timeout 300 python -m pytest tests/test_gpu_dft.py -x -q > gpurun_out/t_dft.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_dft.log
