#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --workload eval --steps 10 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/bench_eval_${TAG}.json 2> gpurun_out/bench_eval.err; echo "rc=$?"; cat gpurun_out/bench_eval_${TAG}.json; tail -3 gpurun_out/bench_eval.err
$CMD > gpurun_out/plain_eval.log 2>&1 &&
for K in spec_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
