#!/bin/bash
# Round-2 ncu evidence: (1) launch list of the bench command, (2) --set full captures of the product kernels.
# Each ncu run only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_ncu_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/plain_bench2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'dft_fold3m_fwd|dft_fold3_fwd|prepare_bulk|synth_train_kernel|synth_fwd_kernel|spec_prep_vec|idft_fold2|bss_corr|bss_solve' \
  -c 14 -f -o gpurun_out/r2_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
