#!/bin/bash
# usage: tools/gpu_round.sh <tag>   -- tests + smoke + gemm sweep + bench (+reference arm)
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
echo "== gpu tests =="
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log
echo "== gemm sweep =="
timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_${TAG}.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_sweep_${TAG}.log
echo "== bench =="
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench.err
