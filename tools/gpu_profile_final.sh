#!/bin/bash
# Round-end capture: tests, smoke, train + eval bench lines, reference arm, ncu launch list, one --set full per kernel.
set -u
mkdir -p gpurun_out
R=${1:-r1z}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${R}.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_${R}.json
timeout 600 python bench.py --workload eval --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eval_${R}.json 2>> gpurun_out/bench.err; echo "eval rc=$?"; cat gpurun_out/bench_eval_${R}.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${R}.json 2>> gpurun_out/bench.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_${R}.json
timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_${R}.log 2>&1; cat gpurun_out/gemm_sweep_${R}.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in dft_fold3_fwd_kernel synth_train_kernel prepare_vec_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${R}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
CMDE="python bench.py --workload eval --steps 2 --warmup 1 --no-cpu-baseline"
$CMDE > gpurun_out/plain_eval.log 2>&1 &&
for K in dft_gemm2_kernel spec_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${R}_$K $CMDE > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
