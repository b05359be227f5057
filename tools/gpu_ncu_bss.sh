#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/time_bss.py > gpurun_out/plain_bss.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bss_' -c 24 --csv --log-file gpurun_out/bss_launches.csv python tools/time_bss.py > gpurun_out/ncu_bss.log 2>&1
echo rc=$?
