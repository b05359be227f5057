#!/bin/bash
set -u
mkdir -p gpurun_out
export AVSEP_DFT_FUSED_STAGE=1 AVSEP_DFT_FOLD3_GROUPS=${SETS:-1} AVSEP_DFT_FOLD_PREFETCH=1
timeout 90 python tools/time_stft.py 2>&1 | tail -1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dft_fold3 -s 6 -c 1 -o gpurun_out/fold3 -f python tools/time_stft.py > gpurun_out/ncu_fold3.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fold3.log
ls -la gpurun_out/fold3.ncu-rep
