#!/usr/bin/env python
"""Time avsep_stft alone (240 signals of 65535 samples) (the folded kernels; MIX=1: two-source mixtures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402

S, L = int(os.environ.get("S", 240)), 65535
w = torch.randn(S, L, device="cuda") * 0.1
plan = av.Plan(S, L)
mag = torch.empty(S, 512, 256, device="cuda")
lib, chk = av._lib.lib, av._lib.check
st = torch.cuda.current_stream().cuda_stream
if os.environ.get("MIX"):                        # two sources per mixture, mixture spectrum = X0 + X1 (dft_fold3m)
    Bm = S // 3
    src = w[:2 * Bm].reshape(2, Bm, L)
    call = lambda: chk(lib.avsep_stft_mix(plan._h, src.data_ptr(), Bm, 2, mag.data_ptr(), None, st))
else:
    call = lambda: chk(lib.avsep_stft(plan._h, w.data_ptr(), S, mag.data_ptr(), None, None, None, st))
for _ in range(5):
    call()
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print("stft ms %.4f" % best)
