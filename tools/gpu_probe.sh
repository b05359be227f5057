#!/bin/bash
# one-GPU probe call: HBM ceilings by access mix + the SM-partitioned pipeline probe
set -u
mkdir -p gpurun_out
timeout 120 python tools/hbm_probe.py > gpurun_out/hbm_probe.json 2> gpurun_out/hbm_probe.err; echo "hbm rc=$?"; cat gpurun_out/hbm_probe.json
timeout 300 python tools/pipeline_probe.py > gpurun_out/pipeline_probe.json 2> gpurun_out/pipeline_probe.err; echo "pipe rc=$?"
tail -3 gpurun_out/pipeline_probe.err; cat gpurun_out/pipeline_probe.json
