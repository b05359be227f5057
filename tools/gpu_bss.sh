#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -x -q -k "calc_metrics" 2>&1 | tail -3
timeout 600 python tools/time_bss.py > gpurun_out/time_bss.json 2> gpurun_out/time_bss.err; echo "rc=$?"; cat gpurun_out/time_bss.json; tail -3 gpurun_out/time_bss.err
