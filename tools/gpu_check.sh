#!/bin/bash
# One gpurun call: GPU parity tests, smoke, short bench.  Every step is bounded by `timeout` so a
# pipeline bug cannot hold the box.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== dft quick ==" 
timeout 240 python -m pytest tests/test_gpu_dft.py -x -q -s -k "basis or raw_gemm" > gpurun_out/t_dft_quick.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_dft_quick.log
echo "== all gpu tests =="
timeout 900 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
