#!/bin/bash
# ncu --set full capture of the separation-half kernels (run only after the plain command has exited 0)
set -u
mkdir -p gpurun_out
timeout 120 python tools/time_separate.py > gpurun_out/plain_sep.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-idft_fold2_kernel|spec_prep}" -s ${SKIP:-6} -c ${COUNT:-2} \
    -f -o gpurun_out/r2_sep python tools/time_separate.py > gpurun_out/ncu_sep.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu_sep.log
