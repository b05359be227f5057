#!/bin/bash
# Parity of the fused-staging folded GEMM, then a short bench with it on / off.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dft.py -x -q -k "fused_stage or folded_and_plain or ragged or stft_vs_oracle" > gpurun_out/t_fold3.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_fold3.log
for cfg in ${FOLD3_CFGS:-"1 4 1" "1 8 1" "1 4 0" "0 4 1"}; do
  set -- $cfg
  echo "== bench fused=$1 warps=$2 prefetch=$3 =="
  AVSEP_DFT_FUSED_STAGE=$1 AVSEP_DFT_FOLD3_GROUPS=$2 AVSEP_DFT_FOLD_PREFETCH=$3 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f$1_w$2_p$3.json 2> gpurun_out/bench_f$1_w$2_p$3.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_f$1_w$2_p$3.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), [(k["kernel"], round(k["ms"],4), round(k["frac"],3)) for k in d.get("kernels",[])][:3], d["clocks"]["sm_mhz"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_f$1_w$2_p$3.err").read()[-1500:])
PY
done
