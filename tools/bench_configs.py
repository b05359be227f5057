#!/usr/bin/env python
"""Per-config numbers for BASELINE.json's configs other than the bench.py workload (C2), on ONE GPU, device-resident
inputs, CUDA events.  bench.py stays the contract; this is evidence for the parity-test configs:
  C1  B=1,   N=2  eval forward (STFT -> warp -> synth -> sigmoid -> un-warp -> mask -> iSTFT): latency of one mixture
  C3  B=128, N=3  eval forward + SDR/SIR/SAR/SI-SDR of every clip (device evaluator)
  C4  60 s clip = 10 windows of 65535 samples through separate_long_clip (synthesizer masks as the model)
  C5  B in {64 .. 1024}, N=2, K=32 train step (STFT + prepare + synth + loss + backward)
Prints one JSON object."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
L, K = 65535, 32
out = {}


def inputs(B, N, seed):
    waves = bench.make_waves(B, N, seed).to(dev)                     # [N+1, B, L]
    g = torch.Generator().manual_seed(seed + 1)
    feat = torch.randn(B, K, 256, 256, generator=g).to(dev)
    vecs = torch.sigmoid(torch.randn(B, N, K, generator=g)).to(dev)
    return waves, feat, vecs, torch.ones(K, device=dev), torch.zeros(1, device=dev)


# ---- C1 ------------------------------------------------------------------------------------------------
waves, feat, vecs, scale, bias = inputs(1, 2, 1234)
st = av.Step(1, 2, K, L)
wav = torch.empty((1, 3, 65280), device=dev)
ms = bench.time_calls(lambda: st.eval_dev(waves, vecs, feat, scale, bias, wav), 50, warmup=5)
out["C1"] = dict(B=1, N=2, eval_ms=ms, mixtures_per_s=1e3 / ms)
st.close()

# ---- C3 ------------------------------------------------------------------------------------------------
B, N = 128, 3
waves, feat, vecs, scale, bias = inputs(B, N, 1236)
st = av.Step(B, N, K, L)
wav = torch.empty((B, N + 1, 65280), device=dev)
ms_sep = bench.time_calls(lambda: st.eval_dev(waves, vecs, feat, scale, bias, wav), 10, warmup=3)
refs = waves[1:, :, :65280].permute(1, 0, 2)                         # [B, N, L] strided view


def metrics():
    return [av.bss_eval_batch(refs, wav[:, 1:])]


metrics()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = metrics()
torch.cuda.synchronize()
ms_met = (time.perf_counter() - t0) * 1e3
sdr = float(torch.cat([r["sdr"] for r in res]).mean())
out["C3"] = dict(B=B, N=N, separate_ms=ms_sep, metrics_ms=ms_met, clips_per_s_with_metrics=B / (ms_sep + ms_met) * 1e3,
                 mixtures_per_s_separation_only=B / ms_sep * 1e3, mean_sdr_db=sdr)
st.close()
del feat, wav, refs
torch.cuda.empty_cache()

# ---- C4 ------------------------------------------------------------------------------------------------
clip = bench.make_waves(1, 1, 1237)[0, 0]
long_mix = torch.cat([clip] * 11)[:661500].to(dev)                   # 60 s at 11025 Hz
W = 11                                                               # ceil((661500 - 65535) / 65535) + 1 windows
g = torch.Generator().manual_seed(7)
featW = torch.randn(W, K, 256, 256, generator=g).to(dev)
vecW = torch.sigmoid(torch.randn(W, 2, K, generator=g)).to(dev)
sc, bi = torch.ones(K, device=dev), torch.zeros(1, device=dev)


def mask_fn(log_mix, mix_w):
    n = log_mix.shape[0]
    m = av.synth_masks(featW[:n], vecW[:n], sc, bi)                  # [N, W, 1, 256, T]
    return m.reshape(2, n, 256, -1)


def run_long():
    return av.separate_long_clip(long_mix, mask_fn)


ms = bench.time_calls(run_long, 10, warmup=3)
nw = run_long().shape[0]
out["C4"] = dict(samples=661500, windows=nw, ms=ms, seconds_of_audio_per_s=60.0 / (ms * 1e-3))

# ---- C5 ------------------------------------------------------------------------------------------------
out["C5"] = []
for B in (64, 128, 256, 512, 1024):
    waves, feat, vecs, scale, bias = inputs(B, 2, 1238)
    st = av.Step(B, 2, K, L)
    dfeat = torch.empty_like(feat)
    dvecs = torch.empty_like(vecs)
    loss = torch.empty(1, device=dev)
    ms = bench.time_calls(lambda: st.train_dev(waves, vecs, feat, scale, bias, dfeat, dvecs, None, None, loss), 10, warmup=3)
    out["C5"].append(dict(B=B, train_ms=ms, mixtures_per_s=B / ms * 1e3, loss=float(loss)))
    st.close()
    del feat, dfeat, waves
    torch.cuda.empty_cache()

print(json.dumps(out))
