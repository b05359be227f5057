#!/bin/bash
# full GPU test suite + smoke + the default bench line
set -u
mkdir -p gpurun_out
echo "== all gpu tests =="
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench =="
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | tail -3; echo "rc=$?"; head -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
