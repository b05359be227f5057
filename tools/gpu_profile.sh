#!/bin/bash
# One gpurun call: re-run the GPU tests that changed, a bench line, then the ncu launch list and one
# `--set full` capture per hot kernel (each ncu run only after the same command exited 0 without ncu).
set -u
mkdir -p gpurun_out
ROUND=${1:-r1}
echo "== gpu tests =="
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_gpu.log
echo "== bench =="
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== ncu =="
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${ROUND}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in synth_bwd_kernel prepare_kernel dft_gemm_kernel synth_fwd_kernel wave_split_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${ROUND}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la gpurun_out
