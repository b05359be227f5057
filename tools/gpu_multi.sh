#!/bin/bash
# N-GPU evidence run: the driver-contract bench line and the BASELINE config sweep (C4 long clip, C5 batch sweep) under torchrun.
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
  tools/bench_configs.py > gpurun_out/configs_${N}gpu.json 2> gpurun_out/configs_${N}gpu.err; echo "configs rc=$?"
grep -v "NCCL INFO" gpurun_out/configs_${N}gpu.err | tail -3
tail -c 1500 gpurun_out/configs_${N}gpu.json
