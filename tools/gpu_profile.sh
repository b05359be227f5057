#!/bin/bash
# One gpurun call: GPU tests, a bench line, then the ncu launch list and one `--set full` capture per
# requested kernel (each ncu run only after the same command exited 0 without ncu).
#   usage: tools/gpu_profile.sh <round-tag> [kernel-regex ...]
set -u
mkdir -p gpurun_out
ROUND=${1:-r1}; shift || true
KERNELS=${@:-"synth_bwd_kernel prepare_vec_kernel dft_gemm_kernel synth_fwd_kernel wave_split_kernel"}
echo "== gpu tests =="
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${ROUND}.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench_${ROUND}.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm =="
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${ROUND}.json 2>> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench_ref_${ROUND}.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
echo "== ncu =="
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${ROUND}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${ROUND}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
