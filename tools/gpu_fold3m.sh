#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -x -q 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_dft.py -x -q 2>&1 | tail -6
for ml in 1 0; do
  AVSEP_DFT_MIX_LINEAR=$ml timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ml$ml.json 2> gpurun_out/bench_ml$ml.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ml$ml.json").read().strip().splitlines()[-1])
    print("mix_linear=$ml value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), [(k["kernel"], round(k["ms"],4), round(k["frac"],3)) for k in d["kernels"]][:3], d["gpu_launches"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_ml$ml.err").read()[-1500:])
PY
done
