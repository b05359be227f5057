#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/diag_steps.py > gpurun_out/diag_steps.log 2>&1; echo "rc=$?"; cat gpurun_out/diag_steps.log
