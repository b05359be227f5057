#!/bin/bash
set -u
mkdir -p gpurun_out
R=r1z
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 250 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${R}.json 2> gpurun_out/bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph"
timeout 150 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 0; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dft_fold3m_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_${R}_dft_fold3m_fwd_kernel $CMD > gpurun_out/ncu_dft_fold3m_fwd_kernel.log 2>&1
echo "ncu rc=$?"
