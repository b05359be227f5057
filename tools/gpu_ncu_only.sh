#!/bin/bash
# ncu captures only (after the same commands ran clean in gpu_profile_final.sh): launch list + --set full per kernel.
set -u
mkdir -p gpurun_out
R=${1:-r1z}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in dft_fold3_fwd_kernel synth_train_kernel prepare_vec_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${R}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
timeout 300 python bench.py --workload eval --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eval_${R}.json 2>> gpurun_out/bench.err; echo "eval rc=$?"
