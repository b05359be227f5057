#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dft tests =="
timeout 600 python -m pytest tests/test_gpu_dft.py -x -q -k "istft or separate or inverse or roundtrip" > gpurun_out/t_inv.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_inv.log
echo "== time separate =="
timeout 120 python tools/time_separate.py > gpurun_out/time_sep_c2.json 2> gpurun_out/time_sep.err; echo "rc=$?"; cat gpurun_out/time_sep_c2.json; tail -3 gpurun_out/time_sep.err
B=128 N=3 timeout 120 python tools/time_separate.py > gpurun_out/time_sep_c3.json 2>> gpurun_out/time_sep.err; echo "rc=$?"; cat gpurun_out/time_sep_c3.json
