#!/bin/bash
# one-GPU check of the staged inverse-operand kernel: parity test, then timing at the C2 and C3 shapes
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dft.py -x -q -k "staged_operand or separate" > gpurun_out/t_spec.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_spec.log
B=80 N=2 timeout 120 python tools/time_separate.py > gpurun_out/time_sep_c2.json 2> gpurun_out/time_sep.err; cat gpurun_out/time_sep_c2.json
B=128 N=3 timeout 120 python tools/time_separate.py > gpurun_out/time_sep_c3.json 2>> gpurun_out/time_sep.err; cat gpurun_out/time_sep_c3.json
tail -3 gpurun_out/time_sep.err
