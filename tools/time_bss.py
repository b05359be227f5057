#!/usr/bin/env python
"""Time the batched device evaluator (bss_metrics.bss_eval_batch) at config C2/C3 sizes, beside the per-clip host
loop it replaces (the CPU restatement in oracle/bss_eval.py on a few clips -- tools/ only, not a product path)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402
from oracle import bss_eval as BSS  # noqa: E402

out = []
for B, N in ((80, 2), (128, 3)):
    L = 65280
    g = torch.Generator(device="cuda").manual_seed(B)
    ref = torch.randn(B, N, L, device="cuda", generator=g) * 0.1
    est = ref + 0.05 * torch.randn(B, N, L, device="cuda", generator=g) + 0.03 * ref.flip(1)
    run = lambda: [av.bss_eval_batch(ref, est)]                                     # hand-written kernels, one call
    run_lib = lambda: [av.bss_eval_batch_torch(ref[b0:b0 + 32], est[b0:b0 + 32]) for b0 in range(0, B, 32)]   # round-1 library route
    run()
    torch.cuda.synchronize()
    t_gpu = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        res = run()
        torch.cuda.synchronize()
        t_gpu = min(t_gpu, time.perf_counter() - t0)
    run_lib()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res_lib = run_lib()
    torch.cuda.synchronize()
    t_lib = time.perf_counter() - t0
    nclip = 2
    t0 = time.perf_counter()
    cpu = [BSS.bss_eval_sources(ref[b].double().cpu().numpy(), est[b].double().cpu().numpy()) for b in range(nclip)]
    t_cpu = (time.perf_counter() - t0) / nclip
    err = max(float(np.max(np.abs(res[0]["sdr"][b].cpu().numpy() - cpu[b][0]))) for b in range(nclip))
    lib_diff = float((torch.cat([r["sdr"] for r in res_lib]) - res[0]["sdr"]).abs().max())
    out.append(dict(B=B, N=N, L=L, gpu_s_per_batch=t_gpu, gpu_clips_per_s=B / t_gpu, torch_library_s_per_batch=t_lib,
                    torch_library_clips_per_s=B / t_lib, max_abs_sdr_diff_vs_library_db=lib_diff, cpu_s_per_clip=t_cpu,
                    cpu_clips_per_s=1.0 / t_cpu, max_abs_sdr_diff_db=err,
                    peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30))
print(json.dumps(out))
