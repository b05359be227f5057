#!/usr/bin/env python
"""Time K2 (prepare_vec_kernel: warp + derive) alone at the C2 (B=80, N=2) and C3 (B=128, N=3) shapes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av, bench  # noqa: E402
out = {}
for B, N in ((80, 2), (128, 3)):
    mag = torch.rand((N + 1, B, 1, 512, 256), device="cuda") + 0.01
    f = lambda: av.ops.prepare(mag[0], [mag[1 + n] for n in range(N)])
    lib, chk = av._lib.lib, av._lib.check
    outs = [torch.empty((B, 256, 256), device="cuda") for _ in range(3 + 2 * N)]
    mags = av._lib.ptr_array([mag[1 + n].data_ptr() for n in range(N)])
    mw = av._lib.ptr_array([outs[3 + n].data_ptr() for n in range(N)])
    gg = av._lib.ptr_array([outs[3 + N + n].data_ptr() for n in range(N)])
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: chk(lib.avsep_prepare(mag[0].data_ptr(), mags, N, B, 256, 1, 1, 1, outs[0].data_ptr(), mw, outs[1].data_ptr(),
                                         outs[2].data_ptr(), gg, st))
    ms = min(bench.time_calls(call, 20, warmup=3) for _ in range(3))
    a = bench.alg("prepare", N)["bytes"] * B
    out[f"B{B}_N{N}"] = dict(ms=ms, gbs=a / ms / 1e6, frac=a / ms / 1e6 / bench.peaks()["hbm"])
print(json.dumps(out))
