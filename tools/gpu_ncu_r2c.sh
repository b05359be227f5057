#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/time_bss.py > gpurun_out/plain_bss.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bss_corr|bss_solve' -s 8 -c 2 -f -o gpurun_out/r2_full_bss \
  python tools/time_bss.py > gpurun_out/ncu_full_bss.log 2>&1
echo "bss rc=$?"; grep -c PROF gpurun_out/ncu_full_bss.log
