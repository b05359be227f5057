#!/bin/bash
# Round-end capture, every command bounded by `timeout`: GPU tests, smoke, train / eval / reference bench lines, then (only
# after the same command ran clean without ncu) the ncu launch list and one --set full capture per kernel of the step.
set -u
mkdir -p gpurun_out
R=${1:-r1z}
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${R}.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 200 python bench.py --workload eval --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_eval_${R}.json 2>> gpurun_out/bench.err; echo "eval rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${R}.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph"
timeout 200 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 0; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in dft_fold3m_fwd_kernel synth_train_kernel prepare_vec_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${R}_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
