#!/bin/bash
# Round-end evidence in one gpurun call (1 GPU): GPU tests, smoke, the default bench line, then -- each only after the same
# command has exited 0 without ncu -- the ncu launch list of the bench command and `--set full` captures of the product kernels
# of both halves (train C2, eval C3) and of the BSS evaluator.  Every command is bounded by `timeout`.
# gpurun copies back at most 64 MiB, and a --set full report is ~5 MB per kernel: PART=a (tests, smoke, bench lines, launch list,
# train kernels, BSS evaluator) and PART=b (eval kernels) are two calls.
set -u
R=${R:-r2}
PART=${PART:-a}
mkdir -p gpurun_out
if [ "$PART" = "a" ]; then
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/bench.err; echo "reference arm rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_ncu_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
TCMD="python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $TCMD > gpurun_out/plain_bench2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'dft_fold3m_fwd|prepare_bulk_kernel|synth_train_kernel|synth_finalize' \
  -s 8 -c 4 -f -o gpurun_out/${R}_full_train $TCMD > gpurun_out/ncu_full_train.log 2>&1
echo "full train rc=$?"; tail -2 gpurun_out/ncu_full_train.log | cut -c1-160
timeout 300 python tools/time_bss.py > gpurun_out/plain_bss.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bss_corr|bss_solve' -s 8 -c 2 -f -o gpurun_out/${R}_full_bss \
  python tools/time_bss.py > gpurun_out/ncu_full_bss.log 2>&1
echo "bss rc=$?"
fi
if [ "$PART" = "b" ]; then
ECMD="python bench.py --workload eval --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $ECMD > gpurun_out/plain_eval.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'dft_fold3m_fwd|dft_fold3_fwd|prepare_bulk_kernel|synth_fwd_staged|spec_prep_tma|idft_fold2' \
  -s 15 -c 10 -f -o gpurun_out/${R}_full_eval $ECMD > gpurun_out/ncu_full_eval.log 2>&1
echo "full eval rc=$?"; tail -2 gpurun_out/ncu_full_eval.log | cut -c1-160
fi
