#!/bin/bash
# one-GPU check after a kernel change: all GPU parity tests, then the bench with its per-kernel summary
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dft.json 2> gpurun_out/bench_dft.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dft.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'],'pipelined',d['value_pipelined']['value'],'e2e',d['e2e']['value'])
for k in d['kernels']: print(' ',k['kernel'],round(k['ms'],4),round(k['frac'],3))
ev=d['eval']; print('eval',ev['value'],ev['ms_per_step'],'explicit',ev['value_explicit_mixture']['ms_per_step'],'e2e',ev['e2e']['value'],ev['e2e']['ms_per_step'])
for k in ev['kernels']: print(' ',k['kernel'],round(k['ms'],4),round(k['frac'],3))
PY
tail -3 gpurun_out/bench_dft.err
