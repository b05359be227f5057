#!/bin/bash
# Round-2 ncu evidence, part 2: --set full captures of the eval-step kernels (C3 shape) and of the BSS evaluator.
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload eval --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_eval.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'dft_fold3_fwd|prepare_bulk|synth_fwd_kernel|spec_prep_vec|idft_fold2' \
  -s 5 -c 5 -f -o gpurun_out/r2_full_eval $CMD > gpurun_out/ncu_full_eval.log 2>&1
echo "eval rc=$?"; tail -2 gpurun_out/ncu_full_eval.log | cut -c1-200
timeout 300 python tools/time_bss.py > gpurun_out/plain_bss.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bss_corr|bss_solve' -s 16 -c 2 -f -o gpurun_out/r2_full_bss \
  python tools/time_bss.py > gpurun_out/ncu_full_bss.log 2>&1
echo "bss rc=$?"; tail -2 gpurun_out/ncu_full_bss.log | cut -c1-200
