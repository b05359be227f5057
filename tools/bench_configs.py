#!/usr/bin/env python
"""Per-config numbers for BASELINE.json's configs other than the bench.py workloads, device-resident inputs, CUDA events,
max over ranks.  bench.py stays the contract (C2 train line + C3 eval block); this is evidence for the other configs:
  C1  B=1,   N=2  eval forward (STFT -> warp -> synth -> sigmoid -> un-warp -> mask -> iSTFT): latency of one mixture   [rank 0]
  C3  B=128, N=3  eval forward + SDR/SIR/SAR/SI-SDR of every clip (hand-written device evaluator)                      [1 GPU only]
  C4  one 60 s clip in 65535-sample windows through separate_long_clip, windows sharded round-robin over the ranks,
      gathered on every rank (strides 65280 = seamless tiling, and 32768 = overlapping)
  C5  SoP++ synthesizer train step, TOTAL batch in {64 .. 1024} split evenly over the ranks (strong scaling per batch size)
Run plain (1 GPU) or under torchrun (--nproc-per-node 2/4/8).  Rank 0 prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402
import bench  # noqa: E402
from avsep_b200 import parallel  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    parallel.init_from_env("nccl")
import torch.distributed as dist  # noqa: E402

dev = torch.device("cuda", local)
L, K = 65535, 32
out = dict(n_gpus=world)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    barrier()
    return parallel.max_over_ranks(e0.elapsed_time(e1) / iters, device=dev)


def inputs(B, N, seed):
    waves = bench.make_waves(B, N, seed).to(dev)                     # [N+1, B, L]
    g = torch.Generator().manual_seed(seed + 1)
    feat = torch.randn(B, K, 256, 256, generator=g).to(dev)
    vecs = torch.sigmoid(torch.randn(B, N, K, generator=g)).to(dev)
    return waves, feat, vecs, torch.ones(K, device=dev), torch.zeros(1, device=dev)


# ---- C1 (one mixture: rank 0) -------------------------------------------------------------------------------
if rank == 0:
    waves, feat, vecs, scale, bias = inputs(1, 2, 1234)
    st = av.Step(1, 2, K, L)
    st.set_mix_on_device(True)                                       # sources only, as bench.py; graph replay on a side stream
    st.set_graph(True)
    wav = torch.empty((1, 3, 65280), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ms = bench.time_calls(lambda: st.eval_dev(waves[1:], vecs, feat, scale, bias, wav), 50, warmup=5)
    torch.cuda.current_stream().wait_stream(side)
    out["C1"] = dict(B=1, N=2, eval_ms=ms, mixtures_per_s=1e3 / ms)
    st.close()
barrier()

# ---- C3 (bench.py's eval block at N GPUs; here for the single-GPU record) ---------------------------------------
if world == 1:
    B, N = 128, 3
    waves, feat, vecs, scale, bias = inputs(B, N, 1236)
    st = av.Step(B, N, K, L)
    st.set_mix_on_device(True)
    st.set_graph(True)
    wav = torch.empty((B, N + 1, 65280), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ms_sep = bench.time_calls(lambda: st.eval_dev(waves[1:], vecs, feat, scale, bias, wav), 10, warmup=3)
    torch.cuda.current_stream().wait_stream(side)
    refs = waves[1:, :, :65280].permute(1, 0, 2)                     # [B, N, L] strided view
    ms_met = bench.time_calls(lambda: av.bss_eval_batch(refs, wav[:, 1:]), 3, warmup=1)
    sdr = float(av.bss_eval_batch(refs, wav[:, 1:])["sdr"].mean())
    out["C3"] = dict(B=B, N=N, separate_ms=ms_sep, metrics_ms=ms_met, clips_per_s_with_metrics=B / (ms_sep + ms_met) * 1e3,
                     mixtures_per_s_separation_only=B / ms_sep * 1e3, mean_sdr_db=sdr)
    st.close()
    del feat, wav, refs, waves
    torch.cuda.empty_cache()

# ---- C4: one 60 s clip, windows sharded over the ranks --------------------------------------------------------------
clip = bench.make_waves(1, 1, 1237)[0, 0]
long_mix = torch.cat([clip] * 11)[:661500].to(dev)                   # 60 s at 11025 Hz
g = torch.Generator().manual_seed(7)
featW = torch.randn(24, K, 256, 256, generator=g).to(dev)            # stand-in U-Net features of up to 24 windows
vecW = torch.sigmoid(torch.randn(24, 2, K, generator=g)).to(dev)
sc, bi = torch.ones(K, device=dev), torch.zeros(1, device=dev)
out["C4"] = []
for stride in (65280, 32768):
    nwin = av.clip_windows(long_mix, 65535, stride).shape[0]
    mine = list(range(rank, nwin, world))

    def mask_fn(log_mix, mix_w):
        n = log_mix.shape[0]
        m = av.synth_masks(featW[mine[:n]], vecW[mine[:n]], sc, bi)  # [N, W_local, 1, 256, T]
        return m.reshape(2, n, 256, -1)

    def run_long():
        loc = av.separate_long_clip(long_mix, mask_fn, stride=stride, rank=rank, world=world)
        return parallel.gather_windows(loc, nwin, rank, world)

    ms = timed(run_long, 10)
    allw = run_long()
    full = av.assemble_windows(allw[:, 0], stride, 661500)
    covered = min(661500, (nwin - 1) * stride + 65280)
    err = float((full[:covered] - long_mix[:covered]).abs().max())   # the mixture channel round-trips over the whole clip
    out["C4"].append(dict(samples=661500, stride=stride, windows=nwin, windows_per_rank=len(mine), ms=ms,
                          seconds_of_audio_per_s=60.0 / (ms * 1e-3), mixture_roundtrip_max_abs_err=err))
del featW
torch.cuda.empty_cache()

# ---- C5: total batch split over the ranks -----------------------------------------------------------------------------
out["C5"] = []
for B_total in (64, 128, 256, 512, 1024):
    B = B_total // world
    if B < 1:
        continue
    waves, feat, vecs, scale, bias = inputs(B, 2, 1238 + rank)
    st = av.Step(B, 2, K, L)
    st.set_mix_on_device(True)                                       # sources only, mixture spectrum = X0 + X1 (as bench.py)
    st.set_graph(True)
    dfeat = torch.empty_like(feat)
    dvecs = torch.empty_like(vecs)
    loss = torch.empty(1, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ms = timed(lambda: st.train_dev(waves[1:], vecs, feat, scale, bias, dfeat, dvecs, None, None, loss), 10)
    gl = float(parallel.global_mean_loss(loss[0], B)) if world > 1 else float(loss)
    out["C5"].append(dict(B_total=B_total, B_per_gpu=B, train_ms=ms, mixtures_per_s=B_total / ms * 1e3, loss=gl))
    st.close()
    del feat, dfeat, waves
    torch.cuda.empty_cache()

if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
