#!/usr/bin/env python
"""Time the separation half (K4: un-warp + threshold + mask -> inverse DFT GEMM with fused overlap-add) alone,
folded vs plain inverse kernel.  B, N from the environment (default C2 shape family B=80, N=2; C3 is B=128 N=3)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402

B, N, L, T = int(os.environ.get("B", 80)), int(os.environ.get("N", 2)), 65535, 256
S = B * (N + 1)
g = torch.Generator(device="cuda").manual_seed(3)
re = torch.randn((B, 512, T), device="cuda", generator=g)
im = torch.randn((B, 512, T), device="cuda", generator=g)
masks = torch.rand((N, B, 256, T), device="cuda", generator=g)
wav = torch.empty((B, N + 1, 65280), device="cuda")
plan = av.Plan(S, L)
lib, chk = av._lib.lib, av._lib.check
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


sep = lambda: chk(lib.avsep_separate(plan._h, masks.data_ptr(), N, B, None, None, re.data_ptr(), im.data_ptr(), 1, 1, 0.5, 1,
                                     wav.data_ptr(), st))
inv = lambda: chk(lib.avsep_debug_gemm_only(plan._h, 1, S, wav.data_ptr(), st))
out = dict(B=B, N=N, signals=S)
for fold in (1, 0):
    plan.set_inverse_fold(bool(fold))
    t_sep, t_inv = timeit(sep), timeit(inv)
    key = "folded" if fold else "plain"
    out[key] = dict(separate_ms=t_sep, inverse_gemm_ms=t_inv, spec_prep_ms=t_sep - t_inv,
                    algorithmic_tflops=0.536e9 * S / (t_inv * 1e-3) / 1e12)
plan.set_inverse_fold(True)
for staged in (1, 0):                                  # operand kernel: shared-memory staged (TMA stores) vs register tile
    plan.set_spec_staged(bool(staged))
    t_sep, t_inv = timeit(sep), timeit(inv)
    nbytes = B * (N * 256 * T * 4 + 2 * 512 * T * 4 + (N + 1) * T * 1024 * 4)
    out["spec_staged" if staged else "spec_register_tile"] = dict(spec_prep_ms=t_sep - t_inv, gbytes_per_s=nbytes / (t_sep - t_inv) / 1e6)
plan.set_spec_staged(True)
print(json.dumps(out))
