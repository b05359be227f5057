#!/usr/bin/env python
"""Benchmark of the mix-and-separate spectrogram hot path (BASELINE.json: separated mixtures/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload both|train|eval]

Workload (config C2, BASELINE.json configs[1]): MUSIC N=2 training step, batch 80 per GPU:
STFT of the (N+1) signals of every mixture (65535 samples @ 11025 Hz, n_fft 1022 / hop 256) ->
prepare (log-frequency warp 512->256, weights, ground-truth masks, log magnitude) -> K=32
synthesizer masks + sigmoid + weighted BCE -> full backward (dfeat_sound, dvecs, dscale, dbias).
Synthetic MUSIC-shaped inputs (harmonic tones); feat_sound ~ N(0,1) stands in
for the U-Net output and is device-resident, exactly as in the reference.

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = the same step through
the host-buffer C-ABI entry point (pinned waveforms + K-vectors H2D, loss D2H inside the timed region).
The line also carries an `eval` block: config C3 as BASELINE.json states it (B = 128, N = 3 per GPU: STFT -> prepare ->
synth + sigmoid -> un-warp / threshold / mask -> iSTFT, plus the SDR / SIR / SAR / SI-SDR evaluator on the device) with
its own device-resident value, pipelined e2e and per-kernel rooflines.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L = 65535
N_SRC = 2
K_CH = 32
B_PER_GPU = 80
T = 256
HW = 256 * 256
METRIC = "separated mixtures/sec"
UNIT = "mixtures/s"
TRAIN_DEPTH = 3            # train e2e steps in flight (each on its own stream)
EVAL_DEPTH = 3             # eval e2e steps in flight: H2D of step i + 2, kernels of i + 1 and D2H of i overlap
PIPE_SM_BUDGET = 80        # SMs the STFT of the two-steps-in-flight arm may occupy (tools/pipeline_probe.py: best of 48..80)

# algorithmic bytes / flops per mixture (each boundary tensor touched once, fp32) as a function of N
SIG_FLOPS = 0.5357e9                              # one [256,1022] x [1022,1024] real-input DFT (or its inverse), fp32-equivalent
IMG = 256 * 256 * 4                               # one [256,256] fp32 plane
SPEC = 512 * 256 * 4                              # one [512,256] fp32 plane


def alg(name, N):
    """dict(bound, bytes | flops, executed = executed / algorithmic tensor work) of one kernel of the path for N sources."""
    table = {
        # forward transform of the mixture and the N sources; executed: bf16 x 3 products on the folded (half-size)
        # contraction = 1.5 x algorithmic; the two-source training kernel runs 2 of the 3 transforms (1.0 x)
        "stft_fwd": dict(bound="tensor", flops=SIG_FLOPS * (N + 1), executed=1.5),
        "prepare": dict(bound="hbm", bytes=(N + 1) * SPEC + (N + 1) * IMG + (N + 2) * IMG),     # warp in + warped out + log, weight, N gt
        # fused forward + loss + backward: feat_sound in (once), dfeat out, masks out, gt + weight in, K-vectors
        "synth_train": dict(bound="hbm", bytes=2 * 32 * IMG + N * IMG + (N + 1) * IMG + N * 32 * 4),
        "synth_fwd_loss": dict(bound="hbm", bytes=32 * IMG + N * IMG + (N + 1) * IMG + N * 32 * 4),
        "synth_bwd": dict(bound="hbm", bytes=32 * IMG + (2 * N + 1) * IMG + 32 * IMG),
        # un-warp in + (mix re, im) in -> bf16 hi/lo operand of the inverse GEMM out ((N+1) signals x 256 x 1024 x 2 x 2 B)
        "spec_prep": dict(bound="hbm", bytes=N * IMG + 2 * SPEC + (N + 1) * 256 * 1024 * 4),
        # inverse transform, folded: bf16 x 3 on two 512 x 512 contractions = 1.5 x algorithmic (the plain kernel: 3 x)
        "idft_inv": dict(bound="tensor", flops=SIG_FLOPS * (N + 1), executed=1.5),
    }
    return table[name]


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel, from the committed `ncu --set full` summaries
    (profiles/r2_ncu_full_summary.txt, then r1_..., produced by tools/ncu_summary.py); None if the kernel is in neither."""
    for name in ("r2_ncu_full_summary.txt", "r1_ncu_full_summary.txt"):
        t = _ncu_traffic_in(os.path.join(ROOT, "profiles", name), kernel_substr)
        if t is not None:
            return t
    return None


def _ncu_traffic_in(path, kernel_substr):
    if not os.path.exists(path):
        return None
    tot, inside, seen = 0.0, False, 0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(path):
        if line.startswith("kernel:"):
            if inside and seen:
                break
            inside = kernel_substr in line
        elif inside and ("dram__bytes_read.sum " in line or "dram__bytes_write.sum " in line):
            parts = line.split()
            if len(parts) >= 3 and parts[-1] in scale:
                tot += float(parts[-2]) * scale[parts[-1]]
                seen += 1
    return tot if seen == 2 else None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p["bf16_tflops"]),
                    tensor_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


def make_waves(B, N, seed):
    """Harmonic-tone sources [N+1, B, L] (index 0 = mixture = sum of sources / already divided by N)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float64) / 11025.0
    f0 = torch.exp(torch.empty(B, N, dtype=torch.float64).uniform_(np.log(110.0), np.log(880.0), generator=g))
    sig = torch.zeros(B, N, L, dtype=torch.float64)
    for h in range(1, 9):
        ph = torch.empty(B, N, 1, dtype=torch.float64).uniform_(0, 2 * np.pi, generator=g)
        sig += (1.0 / h) * torch.sin(2 * np.pi * h * f0[..., None] * t + ph)
    am = torch.empty(B, N, 1, dtype=torch.float64).uniform_(0.2, 2.0, generator=g)
    sig *= 0.5 + 0.5 * torch.sin(2 * np.pi * am * t)
    sig += 0.01 * torch.randn(B, N, L, dtype=torch.float64, generator=g)
    peak = torch.empty(B, N, 1, dtype=torch.float64).uniform_(0.5, 1.0, generator=g)
    sig *= peak / sig.abs().amax(dim=-1, keepdim=True)
    sig = (sig.clamp_(-1, 1) / N).float()
    src = sig.permute(1, 0, 2).contiguous()                       # [N, B, L]
    mix = src.sum(dim=0, keepdim=True)
    return torch.cat([mix, src], dim=0).contiguous()              # [N+1, B, L]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def lines(self):
        try:
            return sum(1 for r in open(self.path) if r.strip())
        except Exception:
            return 0

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            rows = [r for r in rows if len(r) >= 10]
            loaded = [r for r in rows if float(r[9]) >= 50.0] or rows       # samples taken while the GPU was busy
            total = len(rows)
            rows = loaded
            sm = [float(r[1]) for r in rows]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.strip().lower().startswith("active")})
            if sm:
                out = dict(sm_mhz=statistics.median(sm), sm_max_mhz=float(rows[0][2]), reasons=reasons, samples=total,
                           samples_under_load=len(sm), power_w_max=max(float(r[3]) for r in rows))
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def time_calls(fn, iters, warmup=2):
    """Average device time (ms) of fn() over `iters` calls, CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's own torch path on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(B, seed):
    from oracle import torch_ref_path as TP
    waves = make_waves(B, N_SRC, seed)                            # [N+1, B, L]
    waves_bsl = waves.permute(1, 0, 2).contiguous()               # [B, N+1, L]
    g = torch.Generator().manual_seed(seed + 1)
    feat = torch.randn(B, K_CH, 256, 256, generator=g)
    vecs = torch.sigmoid(torch.randn(B, N_SRC, K_CH, generator=g))
    scale = torch.ones(K_CH, requires_grad=True)
    bias = torch.zeros(1, requires_grad=True)

    def step():
        f = feat.clone().requires_grad_(True)
        v = vecs.clone().requires_grad_(True)
        scale.grad = None
        bias.grad = None
        return TP.train_step(waves_bsl, f, v, scale, bias)
    return step


def run_cpu_baseline(B, reps, budget_s=30.0):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_factory(B, 4242)
    step()                                                        # warm-up (first-shape FFT planning etc.)
    times = []
    t_start = time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    return B / min(times), times


def main_reference(args, rank, world):
    if rank != 0:
        return
    B = 16
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_reference_step_factory(B, 4242)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    cores = torch.get_num_threads()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="fp32", data="synthetic", impl="reference",
                config=dict(workload="C2 MUSIC N=2 train step (STFT+prepare+synth K=32+sigmoid+weighted BCE+backward), reference torch "
                                     "path on host cores", batch_per_step=B, num_mix=N_SRC, K=K_CH, stft="1022/256", grid="256x256"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{B} mixtures per step x {args.steps} steps of the C2 workload; torch.stft stands in for librosa "
                                         f"(absent); reference InnerProd/BCELoss call sequence restated in oracle/torch_ref_path.py"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def kernel_entry(name, N, B, ms, pk, share=None, executed=None, note=None):
    """One `kernels[]` entry: achieved = algorithmic bytes (or FLOPs) per launch / measured duration."""
    a = alg(name, N)
    if a["bound"] == "tensor":
        ach = a["flops"] * B / (ms * 1e-3) / 1e12
        ex = a["executed"] if executed is None else executed
        d = dict(kernel=name, ms=ms, share=share, bound="tensor", achieved=ach, peak=pk["tensor"], unit="TFLOP/s",
                 frac=ach / pk["tensor"], executed_over_algorithmic=ex, executed_tflops=ex * ach,
                 tensor_pipe_frac=ex * ach / pk["tensor"], frac_of_sustained_peak=ach / pk["tensor_sustained"],
                 frac_of_nominal_2250=ach / 2250.0, algorithmic_flops_per_launch=a["flops"] * B)
    else:
        ach = a["bytes"] * B / (ms * 1e-3) / 1e9
        d = dict(kernel=name, ms=ms, share=share, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                 frac_of_nominal_8000=ach / 8000.0, algorithmic_bytes_per_launch=a["bytes"] * B)
    if note:
        d["note"] = note
    return d


def roofline_of(top, ncu_name, pk, note):
    return dict(kernel=top["kernel"], bound=top["bound"], achieved=top["achieved"], peak=top["peak"], unit=top["unit"],
                frac=top["frac"], frac_of_nominal=top.get("frac_of_nominal_8000", top.get("frac_of_nominal_2250")),
                peak_nominal=(8000.0 if top["bound"] == "hbm" else 2250.0), traffic=ncu_traffic(ncu_name),
                traffic_unit="bytes per launch (dram read + write), ncu --set full: profiles/r2_ncu_full_summary.txt (r1_... for "
                             "kernels unchanged since round 1)",
                algorithmic_bytes_per_launch=top.get("algorithmic_bytes_per_launch"),
                algorithmic_flops_per_launch=top.get("algorithmic_flops_per_launch"), peak_source=pk["source"], note=note)


class Reducer:
    """The path's only collective: one NCCL all-reduce of a few scalars per step, issued
    asynchronously so it overlaps the next step (NCCL runs it on its own stream after the step's kernels)."""

    def __init__(self, dist, world, dev, width, slots=8):
        self.dist, self.world = dist, world
        self.buf = torch.zeros((slots, width), device=dev)
        self.staged = torch.zeros((slots, width), device=dev)      # per-slot input: [src..., weight], the weight column written once
        self.work = [None] * slots
        self.i = 0
        self.count = 0
        self.weight = None

    def push(self, src, weight):
        """Reduce [src..., weight] of this step; src is a small device tensor the step has just written."""
        if self.world == 1:
            return
        if self.weight != weight:                                  # (re)write the constant weight column of every slot
            self.staged[:, src.numel():] = float(weight)
            self.weight = weight
        k = self.i % len(self.work)
        if self.work[k] is not None:
            self.work[k].wait()                                    # stream-level wait, the host is not blocked
        slot = self.buf[k]
        slot.copy_(self.staged[k], non_blocking=True)              # D2D memcpys on the step's stream: the weight column ...
        slot[:src.numel()].copy_(src.reshape(-1), non_blocking=True)          # ... and the step's own sums
        self.work[k] = self.dist.all_reduce(slot, async_op=True)
        self.i += 1
        self.count += 1

    def push_packed(self, packed):
        """Same, for a step that writes its sums straight into packed[:-1] with packed[-1] = the (constant) weight: one D2D
        copy per step instead of two."""
        if self.world == 1:
            return
        k = self.i % len(self.work)
        if self.work[k] is not None:
            self.work[k].wait()
        self.buf[k].copy_(packed, non_blocking=True)
        self.work[k] = self.dist.all_reduce(self.buf[k], async_op=True)
        self.i += 1
        self.count += 1

    def finish(self):
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None] * len(self.work)
        torch.cuda.synchronize()
        k = (self.i - 1) % len(self.work) if self.i else 0
        return self.buf[k].tolist()


def make_pcm(B, N, seed):
    """The same harmonic tones as 16-bit PCM [N, B, L] int16 (what a wav file holds), for the opt-in PCM hand-off."""
    w = make_waves(B, N, seed)[1:] * N                            # undo the / N: full-scale sources in [-1, 1]
    return torch.clamp(torch.round(w * 32767.0), -32768, 32767).to(torch.int16).contiguous()


def train_section(av, parallel, dist, args, rank, world, local, dev, pk, sampler):
    B, N, K = B_PER_GPU, N_SRC, K_CH
    waves_host = make_waves(B, N, 1234 + rank).pin_memory()                       # [N+1, B, L]
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    feat = torch.randn((B, K, 256, 256), device=dev, generator=g)
    vecs_host = torch.sigmoid(torch.randn((B, N, K), generator=torch.Generator().manual_seed(7 + rank))).pin_memory()
    scale = torch.ones(K, device=dev)
    bias = torch.zeros(1, device=dev)
    waves_dev = waves_host.to(dev)
    vecs_dev = vecs_host.to(dev)
    dfeat = torch.empty_like(feat)
    dvecs = torch.empty((B, N, K), device=dev)
    dscale = torch.empty(K, device=dev)
    dbias = torch.empty(1, device=dev)
    loss_count = torch.tensor([0.0, float(B)], device=dev)         # [loss, count]: the step writes its loss into element 0, and the
    loss = loss_count[:1]                                          # pair is what the per-step all-reduce sends
    dvecs_host = torch.empty((B, N, K)).pin_memory()
    step = av.Step(B, N, K, L)
    step.set_graph(not args.no_graph)                              # train step replayed as a CUDA graph (side streams only)
    src_only = not args.explicit_mixture
    if src_only:
        step.set_mix_on_device(True)
    # Train step: the step is handed the SOURCE waveforms only and forms the mixture on the GPU, as the e2e arm does
    # (dataset/music.py:127 defines the mixture as the sum of its sources); --explicit-mixture hands over the
    # mixture waveform as a third signal instead.
    sources_dev = waves_dev[1:]                                    # [N, B, L] view

    def step_dev():
        if src_only:
            step.train_dev(sources_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)
        else:
            step.train_dev(waves_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ------------------------------------------------------------
    # The clock sampler (nvidia-smi -lms 200) runs from before the warm-up to after the e2e loop so its samples
    # bracket both timed regions.  Its start-up (NVML attach) perturbs kernel launches for a few hundred ms, so the
    # warm-up keeps the GPU loaded until the sampler has produced its first samples (bounded at 4 s).
    t_w = time.perf_counter()
    n_w = 0
    while True:
        step_dev()
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize()
            el = time.perf_counter() - t_w
            if n_w >= max(args.warmup, 3) and (rank != 0 or sampler.proc is None or sampler.lines() >= 3 or el > 4.0) and el > 0.3:
                break
    red = Reducer(dist, world, dev, 2)
    for _ in range(3):                                             # warm the collective (NCCL channel setup) outside the timed region
        red.push_packed(loss_count)
    red.finish()
    barrier()
    launches0 = av.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
        red.push_packed(loss_count)                                # [loss, count] all-reduce of THIS step, inside the timed region
    e1.record()
    red_last = red.finish()
    barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    launches = av.launch_count() - launches0
    loss_val = float(loss)
    value = B * world * args.steps / (ms_total * 1e-3)
    loss_global = (red_last[0] / world) if world > 1 else loss_val  # equal per-rank counts: mean of the per-rank means

    # The same K steps with the mixture handed over as an explicit third waveform (three transforms instead of two),
    # reported beside `value` so both readings of "inputs resident in HBM" are on record.
    explicit = None
    if src_only:
        step.set_mix_on_device(False)
        for _ in range(max(args.warmup, 3)):
            step.train_dev(waves_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(args.steps):
            step.train_dev(waves_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)
        x1.record()
        barrier()
        ms_x = parallel.max_over_ranks(x0.elapsed_time(x1), device=dev)
        explicit = dict(value=B * world * args.steps / (ms_x * 1e-3), unit=UNIT, ms_per_step=ms_x / args.steps, loss=float(loss),
                        note="mixture waveform resident in HBM as a third signal: three STFTs per mixture")
        step.set_mix_on_device(True)

    # ---- end to end through the host-buffer C-ABI entry point (`e2e`) ---------------------------------------
    # Every step copies its pinned host inputs H2D and reads its loss (+ dvecs) back D2H inside the timed
    # region.  TRAIN_DEPTH Step objects alternate on as many streams (a software pipeline, like a prefetching
    # training loop), so the copy of step i+1 overlaps the kernels of step i; the loss of step i is read on
    # the host before its Step object is reused for step i + TRAIN_DEPTH.  (Depth 3 rather than 2: with two steps in flight the
    # host's wake-up + enqueue after every synchronize sits on the critical path of the faster PCM16 arm.)  At N > 1 the
    # [loss, count] all-reduce of every step is issued (asynchronously) inside the loop as well.
    more_steps = [av.Step(B, N, K, L) for _ in range(TRAIN_DEPTH - 1)]
    for s_ in more_steps:
        s_.set_graph(not args.no_graph)
    pipes = [dict(step=s_, stream=torch.cuda.Stream(), dfeat=(dfeat if s_ is step else torch.empty_like(feat)),
                  loss=torch.zeros(1).pin_memory(), dvecs=torch.empty((B, N, K)).pin_memory()) for s_ in [step] + more_steps]
    for p in pipes:
        p["loss_dev"] = p["step"].view("loss")
    sources_host = waves_host[1:]                                  # [N, B, L] view of the same pinned buffer
    pcm_host = make_pcm(B, N, 1234 + rank).pin_memory()            # [N, B, L] int16
    red_e = Reducer(dist, world, dev, 2)

    def run_e2e(k, mode):
        for p in pipes:
            p["step"].set_mix_on_device(mode != "full")
            # the PCM16 arm is compute-bound (0.39 ms of copy per step), so the kernels of the steps in flight overlap: cap the
            # persistent STFT grid as StepPipeline does, otherwise two all-SM STFTs and a synthesizer kernel interleave at random
            p["step"].plan.set_sm_budget(PIPE_SM_BUDGET if mode == "pcm16" else 0)
        for i in range(k):
            p = pipes[i % TRAIN_DEPTH]
            p["stream"].synchronize()                              # step i - TRAIN_DEPTH done: its loss is on the host now
            _ = float(p["loss"][0])
            with torch.cuda.stream(p["stream"]):
                if mode == "pcm16":
                    p["step"].train_pcm16_async(pcm_host, None, vecs_host, feat, scale, bias, p["dfeat"], p["loss"], p["dvecs"])
                else:
                    src = sources_host if mode == "sources" else waves_host
                    p["step"].train_host_async(src, vecs_host, feat, scale, bias, p["dfeat"], p["loss"], p["dvecs"])
                red_e.push(p["loss_dev"], B)
        for p in pipes:
            p["stream"].synchronize()
        red_e.finish()
        return float(pipes[(k - 1) % TRAIN_DEPTH]["loss"][0])

    def timed_e2e(mode):
        torch.cuda.synchronize()
        run_e2e(2 * TRAIN_DEPTH, mode)
        barrier()
        t0 = time.perf_counter()
        lh = run_e2e(args.steps, mode)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return parallel.max_over_ranks(wall, device=dev), lh

    # ---- device-resident, two steps in flight (overlap the tensor-bound STFT with the HBM-bound synthesizer) ----
    # The STFT + prepare front end depends on the waveforms only, not on the model, so batch i + 1's STFT may run beside batch
    # i's synthesizer kernels -- the overlap the reference gets from its DataLoader workers (dataset/base.py:142-147 vs
    # main.py:104-139).  Two Step objects alternate on two streams, inputs resident in HBM, no host synchronisation inside the
    # loop; the STFT's persistent grid is capped at PIPE_SM_BUDGET SMs (avsep_plan_set_sm_budget) so the other step's
    # synthesizer CTAs find free SMs while it runs.  Every step's kernels are the serial step's (bit-identical results);
    # reported beside `value`, which stays the one-step-at-a-time number.
    pipelined = None
    if src_only:
        pipe = av.StepPipeline(B, N, K, L, depth=2, stft_sm_budget=PIPE_SM_BUDGET, graph=not args.no_graph)
        outs = [dict(dfeat=pipes[i]["dfeat"], dv=torch.empty_like(dvecs), ds=torch.empty_like(dscale), db=torch.empty_like(dbias),
                     ls=torch.zeros(1, device=dev)) for i in range(2)]

        def run_pipe(k):
            for _ in range(k):
                slot = pipe.next()
                o = outs[slot.index]
                with torch.cuda.stream(slot.stream):
                    slot.step.train_dev(sources_dev, vecs_dev, feat, scale, bias, o["dfeat"], o["dv"], o["ds"], o["db"], o["ls"])
            pipe.synchronize()

        run_pipe(max(args.warmup, 4))
        barrier()
        y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y0.record()
        run_pipe(args.steps)
        y1.record()
        barrier()
        ms_p = parallel.max_over_ranks(y0.elapsed_time(y1), device=dev)
        pipelined = dict(value=B * world * args.steps / (ms_p * 1e-3), unit=UNIT, ms_per_step=ms_p / args.steps,
                         loss=[float(outs[0]["ls"]), float(outs[1]["ls"])], steps_in_flight=2, stft_sm_budget=PIPE_SM_BUDGET,
                         api="avsep_b200.StepPipeline (two Step objects on two streams, round-robin)",
                         note="two steps in flight on two streams, inputs resident in HBM, whole loop timed with CUDA events (fill and "
                              "drain included; this arm issues no collective); same kernels and bits as `value`; the gain is small because "
                              "the STFT alone already runs the part into its power cap and the synthesizer kernels lose throughput on "
                              "fewer SMs (tools/pipeline_probe.py, profiles/r2_pipeline_probe.json)")
        pipe.close()

    # primary: the host supplies the N source waveforms (fp32), the mixture (their sum, dataset/music.py:127) is formed
    # on the GPU; secondary: the host also supplies the mixture waveform; opt-in: the sources as 16-bit PCM
    ms_e2e, loss_host = timed_e2e("sources")
    ms_e2e_full, loss_host_full = timed_e2e("full")
    ms_e2e_pcm, loss_host_pcm = timed_e2e("pcm16")

    # copy roof: the same pinned buffers through cudaMemcpyAsync alone (no kernels), all ranks at once -- what the
    # host platform can deliver to this many GPUs concurrently; e2e is reported as a fraction of it
    def copy_only(src_t, k):
        dst = torch.empty(src_t.shape, dtype=src_t.dtype, device=dev)
        for _ in range(3):
            dst.copy_(src_t, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            dst.copy_(src_t, non_blocking=True)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return parallel.max_over_ranks(wall, device=dev) / k

    ms_copy_f32 = copy_only(sources_host, args.steps)
    ms_copy_pcm = copy_only(pcm_host, args.steps)

    for p in pipes:
        p["step"].set_mix_on_device(False)
        p["step"].plan.set_sm_budget(0)
    step.set_mix_on_device(src_only)                 # pipes[0] is the value arm's Step: back to the input layout step_dev() passes
    # same inputs, same kernels through the host-buffer entry point: the losses must agree (a gross mismatch = a bug;
    # the exact values are reported in the JSON line)
    assert abs(loss_host - loss_val) <= 1e-3 * max(abs(loss_val), 1.0), (loss_host, loss_val)
    # explicit mixture waveform vs |X0 + X1|: equal to fp32 rounding of the magnitudes (a few binary targets may flip)
    assert abs(loss_host_full - loss_val) <= 1e-2 * max(abs(loss_val), 1.0), (loss_host_full, loss_val)
    # 16-bit quantisation of the same tones: a different (coarser) input, close but not equal
    assert abs(loss_host_pcm - loss_val) <= 5e-2 * max(abs(loss_val), 1.0), (loss_host_pcm, loss_val)
    # keep the GPU running the same step until the 200 ms sampler has seen the loaded state a few times (the timed regions
    # themselves are ~10 ms): the clocks / throttle reasons reported are those of this workload under sustained load
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < 1.2:
        for _ in range(50):
            step_dev()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    per = lambda ms: B * world * args.steps / (ms * 1e-3)
    h2d_src = sources_host.numel() * 4 + vecs_host.numel() * 4
    h2d_full = waves_host.numel() * 4 + vecs_host.numel() * 4
    h2d_pcm = pcm_host.numel() * 2 + vecs_host.numel() * 4
    d2h = 4 + dvecs_host.numel() * 4
    e2e = dict(value=per(ms_e2e), unit=UNIT, h2d_bytes_per_step=h2d_src, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps,
               pipeline_depth=TRAIN_DEPTH, timer="host wall clock around the K steps, synchronised on both sides",
               api="avsep_step_train_host_async (C-ABI): pinned host SOURCE waveforms [N,B,L] fp32 + K-vectors in, mixture summed "
                   f"on the GPU (avsep_step_set_mix_on_device), loss + dvecs out; {TRAIN_DEPTH} steps in flight on {TRAIN_DEPTH} streams"
                   + ("; [loss, count] all-reduced (NCCL, async) every step" if world > 1 else ""),
               copy_roof=dict(ms_per_step=ms_copy_f32, gbytes_per_s_per_gpu=sources_host.numel() * 4 / (ms_copy_f32 * 1e-3) / 1e9,
                              frac=ms_copy_f32 / (ms_e2e / args.steps),
                              note="the same pinned [N,B,L] fp32 buffer through cudaMemcpyAsync alone, every rank of this run copying "
                                   "at once, no kernels: the H2D time the host platform imposes per step; frac = copy-only time / e2e "
                                   "time per step (1.0 = the step is hidden completely behind its own input copy)"),
               with_mixture_h2d=dict(value=per(ms_e2e_full), ms_per_step=ms_e2e_full / args.steps, h2d_bytes_per_step=h2d_full,
                                     note="same call with the mixture waveform also supplied by the host, [N+1,B,L]"),
               pcm16_handoff=dict(value=per(ms_e2e_pcm), ms_per_step=ms_e2e_pcm / args.steps, h2d_bytes_per_step=h2d_pcm,
                                  copy_roof_ms_per_step=ms_copy_pcm, copy_roof_frac=ms_copy_pcm / (ms_e2e_pcm / args.steps),
                                  loss=loss_host_pcm,
                                  note="opt-in avsep_step_train_pcm16_async: the sources as the int16 samples a wav file holds; "
                                       "int16/32768 -> *gain -> clip -> /N on the device in the loader's fp32 order "
                                       "(dataset/base.py:156-172, music.py:120), bit-identical to the fp32 hand-off of the same samples "
                                       "(tests/test_gpu_step.py); not the headline e2e because it changes the loader contract"))

    # ---- per-kernel roofline (rank 0), each kernel timed alone with CUDA events ---------------------------
    kernels, roofline = [], None
    if rank == 0:
        lib, stream, chk = av._lib.lib, torch.cuda.current_stream().cuda_stream, av._lib.check
        plan = lib.avsep_step_plan(step._h)
        S = (N + 1) * B
        mag = torch.empty((S, 512, T), device=dev)
        gt = step.buffer("gt_masks")
        wts = step.buffer("weights")
        pred = torch.empty((N, B, HW), device=dev)
        scratch = torch.empty(int(lib.avsep_synth_scratch_bytes(B, N, K, HW)), dtype=torch.uint8, device=dev)
        one = torch.ones(1, device=dev)
        mags = av._lib.ptr_array([mag.data_ptr() + (1 + n) * B * 512 * T * 4 for n in range(N)])
        outs = [torch.empty((B, 256, T), device=dev) for _ in range(3 + 2 * N)]
        mw = av._lib.ptr_array([outs[3 + n].data_ptr() for n in range(N)])
        gg = av._lib.ptr_array([outs[3 + N + n].data_ptr() for n in range(N)])
        if src_only:                                 # one kernel: two source transforms, mixture spectrum = their sum
            t_stft = time_calls(lambda: chk(lib.avsep_stft_mix(plan, sources_dev.data_ptr(), B, N, mag.data_ptr(), None, stream)), 10)
        else:
            t_stft = time_calls(lambda: chk(lib.avsep_stft(plan, waves_dev.data_ptr(), S, mag.data_ptr(), None, None, None, stream)), 10)
        t_prep = time_calls(lambda: chk(lib.avsep_prepare(mag.data_ptr(), mags, N, B, T, 1, 1, 1, outs[0].data_ptr(), mw,
                                                          outs[1].data_ptr(), outs[2].data_ptr(), gg, stream)), 10)
        t_fwd = time_calls(lambda: chk(lib.avsep_synth_forward(feat.data_ptr(), vecs_dev.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                                               B, N, K, HW, 1, pred.data_ptr(), 3, gt.data_ptr(), wts.data_ptr(),
                                                               loss.data_ptr(), scratch.data_ptr(), stream)), 10)
        t_bwd = time_calls(lambda: chk(lib.avsep_synth_backward(feat.data_ptr(), vecs_dev.data_ptr(), scale.data_ptr(), B, N, K, HW, 1, 3,
                                                                pred.data_ptr(), gt.data_ptr(), wts.data_ptr(), one.data_ptr(), None,
                                                                dfeat.data_ptr(), dvecs.data_ptr(), dscale.data_ptr(), dbias.data_ptr(),
                                                                scratch.data_ptr(), stream)), 10)
        t_train = time_calls(lambda: chk(lib.avsep_synth_train(feat.data_ptr(), vecs_dev.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                                               B, N, K, HW, 1, 3, gt.data_ptr(), wts.data_ptr(), None, pred.data_ptr(),
                                                               loss.data_ptr(), dfeat.data_ptr(), dvecs.data_ptr(), dscale.data_ptr(),
                                                               dbias.data_ptr(), scratch.data_ptr(), stream)), 10)
        times = {"stft_fwd": t_stft, "prepare": t_prep, "synth_train": t_train}
        tot = sum(times.values())
        stft_note = ("algorithmic = fp32-equivalent full DFT of the 3 signals (0.536 GFLOP/signal); the kernel folds / splits the "
                     "waveform itself (no staging kernel) and, since the step forms the mixture from its 2 sources, executes 2 of "
                     "the 3 transforms (mixture spectrum = X0 + X1): bf16 x 3 on the half-size folded contraction x 2/3 = 1.0 x "
                     "algorithmic" if src_only else
                     "algorithmic = fp32-equivalent full DFT (0.536 GFLOP/signal); executed = bf16 x 3 on the folded contraction = 1.5 x")
        for name, ms in times.items():
            kernels.append(kernel_entry(name, N, B, ms, pk, share=ms / tot, executed=(1.0 if src_only else 1.5) if name == "stft_fwd" else None,
                                        note=stft_note if name == "stft_fwd" else None))
        for name, ms in {"synth_fwd_loss": t_fwd, "synth_bwd": t_bwd}.items():
            kernels.append(kernel_entry(name, N, B, ms, pk, note="not part of the step: the unfused forward / backward pair, for comparison"))
        top = max([k for k in kernels if k["share"] is not None], key=lambda k: k["ms"])
        roofline = roofline_of(top, {"synth_train": "synth_train_kernel", "stft_fwd": "dft_fold3m_fwd_kernel<2, 2>",
                                     "prepare": "prepare_bulk_kernel"}[top["kernel"]], pk,
                               "algorithmic fp32-equivalent FLOPs (0.536 GFLOP/signal)" if top["bound"] == "tensor" else
                               "algorithmic bytes per SURVEY.md section 8(d) (feat_sound counted once: fused fwd+bwd)")
    step.close()
    for s_ in more_steps:
        s_.close()
    return dict(value=value, ms_total=ms_total, launches=int(launches), loss=loss_val, loss_global=loss_global, explicit=explicit,
                pipelined=pipelined,
                e2e=e2e, loss_e2e=dict(sources_only=loss_host, explicit_mixture=loss_host_full, pcm16=loss_host_pcm), clocks=clocks,
                kernels=kernels, roofline=roofline, src_only=src_only,
                collective=(dict(per_step="all_reduce([loss, count]) over NCCL, async, inside both timed loops", calls=red.count + red_e.count)
                            if world > 1 else None))


def eval_section(av, parallel, dist, args, rank, world, dev, pk):
    """Config C3 as BASELINE.json states it: MUSIC+ / duet N = 3 mixtures, batch 128 per GPU, eval with iSTFT
    reconstruction and SDR / SIR / SAR (+ SI-SDR)."""
    B, N, K = 128, 3, K_CH
    waves_host = make_waves(B, N, 2234 + rank).pin_memory()                       # [N+1, B, L]
    g = torch.Generator(device=dev).manual_seed(199 + rank)
    feat = torch.randn((B, K, 256, 256), device=dev, generator=g)
    vecs_host = torch.sigmoid(torch.randn((B, N, K), generator=torch.Generator().manual_seed(17 + rank))).pin_memory()
    scale = torch.ones(K, device=dev)
    bias = torch.zeros(1, device=dev)
    waves_dev = waves_host.to(dev)
    vecs_dev = vecs_host.to(dev)
    wav_dev = torch.empty((B, N + 1, 65280), device=dev)
    step = av.Step(B, N, K, L)
    # As in the train arm, the step is handed the N SOURCE waveforms and forms the mixture on the GPU (dataset/music.py:127: the
    # evaluation mixtures are sums of their sources too): the mixture spectrum -- magnitude and the (re, im) the inverse
    # transform consumes -- is the sum of the three source transforms (dft_fold3m<3>), three transforms instead of four.
    # --explicit-mixture hands the mixture waveform over as a fourth signal; that number is reported beside `value` either way.
    src_only = not args.explicit_mixture
    step.set_mix_on_device(src_only)
    step.set_graph(not args.no_graph)                              # eval step replayed as a CUDA graph (side streams only)
    sources_dev, sources_host = waves_dev[1:], waves_host[1:]
    in_dev, in_host = (sources_dev, sources_host) if src_only else (waves_dev, waves_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    red = Reducer(dist, world, dev, 5)
    msum = torch.zeros(4, device=dev)                              # per-step metric sums (filled by the evaluator below)
    explicit = None
    if src_only:                                                   # the explicit-mixture reading first, then the default
        step.set_mix_on_device(False)
        for _ in range(max(args.warmup, 3)):
            step.eval_dev(waves_dev, vecs_dev, feat, scale, bias, wav_dev)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(args.steps):
            step.eval_dev(waves_dev, vecs_dev, feat, scale, bias, wav_dev)
        x1.record()
        barrier()
        ms_x = parallel.max_over_ranks(x0.elapsed_time(x1), device=dev)
        explicit = dict(value=B * world * args.steps / (ms_x * 1e-3), unit=UNIT, ms_per_step=ms_x / args.steps,
                        note="mixture waveform resident in HBM as a fourth signal: four STFTs per mixture")
        step.set_mix_on_device(True)
    for _ in range(max(args.warmup, 3)):
        step.eval_dev(in_dev, vecs_dev, feat, scale, bias, wav_dev)
    for _ in range(2):
        red.push(msum, B)
    red.finish()
    barrier()
    launches0 = av.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step.eval_dev(in_dev, vecs_dev, feat, scale, bias, wav_dev)
        red.push(msum, B)                                          # [sdr, sir, sar, si_sdr sums, clip count] of the step
    e1.record()
    red.finish()
    barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    launches = av.launch_count() - launches0
    value = B * world * args.steps / (ms_total * 1e-3)

    # SDR / SIR / SAR / SI-SDR of every clip on the device (fp64), timed separately: the evaluator, not the
    # separation, bounds the clips/s of an evaluation pass (main.py:252-275 -> asteroid -> mir_eval)
    refs = waves_dev[1:, :, :65280].permute(1, 0, 2)               # [B, N, L] strided view of the source waveforms: scored in place

    def metrics():
        return av.bss_eval_batch(refs, wav_dev[:, 1:])             # hand-written evaluator (csrc/bss_eval.cu), all clips in one call

    metrics()
    barrier()
    t0 = time.perf_counter()
    res = metrics()
    torch.cuda.synchronize()
    ms_met = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, device=dev)
    mean_sdr = float(res["sdr"].mean())
    del res

    # ---- e2e: pinned host waveforms + K-vectors in, separated waveforms out to the host, EVAL_DEPTH steps in flight ----
    # A step is H2D (134 MB) -> kernels -> D2H (134 MB).  With two steps in flight the H2D engine idles while the older step
    # drains its D2H; with three the two DMA engines (one per direction) and the SMs are all kept busy.
    extra = [av.Step(B, N, K, L) for _ in range(EVAL_DEPTH - 1)]
    for s_ in extra:
        s_.set_mix_on_device(src_only)
        s_.set_graph(not args.no_graph)
    pipes = [dict(step=s_, stream=torch.cuda.Stream(), wav=torch.empty((B, N + 1, 65280)).pin_memory()) for s_ in [step] + extra]

    def run_e2e(k, to_host):
        for i in range(k):
            p = pipes[i % EVAL_DEPTH]
            p["stream"].synchronize()                              # step i - EVAL_DEPTH done: its waveforms are on the host now
            with torch.cuda.stream(p["stream"]):
                p["step"].eval_host_async(in_host, vecs_host, feat, scale, bias, p["wav"] if to_host else None)
                red.push(msum, B)
        for p in pipes:
            p["stream"].synchronize()
        red.finish()

    def timed_e2e(to_host):
        torch.cuda.synchronize()
        run_e2e(2 * EVAL_DEPTH, to_host)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps, to_host)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return parallel.max_over_ranks(wall, device=dev)

    ms_e2e = timed_e2e(True)
    ms_e2e_dev = timed_e2e(False)
    per = lambda ms: B * world * args.steps / (ms * 1e-3)
    h2d = in_host.numel() * 4 + vecs_host.numel() * 4
    d2h = B * (N + 1) * 65280 * 4

    kernels, roofline = [], None
    if rank == 0:
        lib, stream, chk = av._lib.lib, torch.cuda.current_stream().cuda_stream, av._lib.check
        plan = lib.avsep_step_plan(step._h)
        S = (N + 1) * B
        mag = torch.empty((S, 512, T), device=dev)
        re = torch.empty((B, 512, T), device=dev)
        im = torch.empty((B, 512, T), device=dev)
        step.eval_dev(in_dev, vecs_dev, feat, scale, bias, wav_dev)
        pred = step.buffer("pred_masks")
        gt = step.buffer("gt_masks")
        wts = step.buffer("weights")
        re.copy_(step.buffer("mix_re")); im.copy_(step.buffer("mix_im"))
        scratch = torch.empty(int(lib.avsep_synth_scratch_bytes(B, N, K, HW)), dtype=torch.uint8, device=dev)
        loss = torch.zeros(1, device=dev)
        mags = av._lib.ptr_array([mag.data_ptr() + (1 + n) * B * 512 * T * 4 for n in range(N)])
        outs = [torch.empty((B, 256, T), device=dev) for _ in range(3 + 2 * N)]
        mw = av._lib.ptr_array([outs[3 + n].data_ptr() for n in range(N)])
        gg = av._lib.ptr_array([outs[3 + N + n].data_ptr() for n in range(N)])
        masks_out = torch.empty((N, B, HW), device=dev)
        if src_only:
            t_stft = time_calls(lambda: chk(lib.avsep_stft_mix_reim(plan, sources_dev.data_ptr(), B, N, mag.data_ptr(), re.data_ptr(), im.data_ptr(), stream)), 10)
        else:
            t_stft = time_calls(lambda: chk(lib.avsep_stft_reim(plan, waves_dev.data_ptr(), S, mag.data_ptr(), re.data_ptr(), im.data_ptr(), B, stream)), 10)
        t_prep = time_calls(lambda: chk(lib.avsep_prepare(mag.data_ptr(), mags, N, B, T, 1, 1, 1, outs[0].data_ptr(), mw,
                                                          outs[1].data_ptr(), outs[2].data_ptr(), gg, stream)), 10)
        t_fwd = time_calls(lambda: chk(lib.avsep_synth_forward(feat.data_ptr(), vecs_dev.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                                               B, N, K, HW, 1, masks_out.data_ptr(), 3, gt.data_ptr(), wts.data_ptr(),
                                                               loss.data_ptr(), scratch.data_ptr(), stream)), 10)
        t_sep = time_calls(lambda: chk(lib.avsep_separate(plan, pred.data_ptr(), N, B, None, None, re.data_ptr(), im.data_ptr(), 1, 1,
                                                          0.5, 1, wav_dev.data_ptr(), stream)), 10)
        t_inv = time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan, 1, S, wav_dev.data_ptr(), stream)), 10)
        times = {"stft_fwd": t_stft, "prepare": t_prep, "synth_fwd_loss": t_fwd, "spec_prep": max(t_sep - t_inv, 1e-6), "idft_inv": t_inv}
        tot = sum(times.values())
        notes = {"stft_fwd": ("folded transform with the operand generated in the kernel, N source transforms for N + 1 spectra (mixture "
                              "spectrum = sum of the source spectra, dft_fold3m<3>): executes 1.5 x N / (N + 1) = 1.125 x algorithmic; also "
                              "writes (re, im) of the mixture for the inverse transform") if src_only else
                             ("folded transform of the N + 1 signals with the operand generated in the kernel (executes 1.5 x algorithmic); "
                              "also writes (re, im) of the mixture for the inverse transform"),
                 "spec_prep": "timed as avsep_separate minus the inverse GEMM alone (the C-ABI has no entry point for it alone)",
                 "idft_inv": "folded inverse GEMM with the overlap-add in the epilogue (idft_fold2_kernel)"}
        for name, ms in times.items():
            kernels.append(kernel_entry(name, N, B, ms, pk, share=ms / tot, note=notes.get(name),
                                        executed=(1.5 * N / (N + 1) if src_only else 1.5) if name == "stft_fwd" else None))
        top = max(kernels, key=lambda k: k["ms"])
        roofline = roofline_of(top, {"stft_fwd": "dft_fold3m_fwd_kernel<2, 3>" if src_only else "dft_fold3_fwd_kernel", "prepare": "prepare_bulk_kernel", "synth_fwd_loss": "synth_fwd_staged_kernel",
                                     "spec_prep": "spec_prep_tma_kernel", "idft_inv": "idft_fold2_kernel"}[top["kernel"]], pk,
                               "largest kernel of the eval step")
    step.close()
    for s_ in extra:
        s_.close()
    return dict(
        config=dict(workload="C3 MUSIC+/duet N=3 eval step: STFT (" + ("3 source transforms, mixture spectrum = X0 + X1 + X2" if src_only else
                             "4 signals") + ", mixture re/im kept)+prepare+synth K=32+sigmoid+BCE+un-warp+"
                             "threshold+mask+iSTFT (overlap-add, envelope, trim, clip), then SDR/SIR/SAR/SI-SDR of every clip on the device",
                    mixture=("formed on the GPU from the 3 device-resident source waveforms (dataset/music.py:127)" if src_only else
                             "explicit mixture waveform (--explicit-mixture)"),
                    batch_per_gpu=B, global_batch=B * world, num_mix=N, K=K, samples=L, stft="1022/256", grid="256x256",
                    l2="inputs larger than L2: 1.07 GB of feat_sound per step vs 126 MB L2"),
        metric=METRIC, unit=UNIT, value=value, ms_per_step=ms_total / args.steps, gpu_launches=int(launches),
        value_explicit_mixture=explicit,
        e2e=dict(value=per(ms_e2e), unit=UNIT, ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 pipeline_depth=EVAL_DEPTH, api="avsep_step_eval_host_async (C-ABI): pinned host " + ("SOURCE waveforms [N,B,L]" if src_only else "waveforms [N+1,B,L]") + " + K-vectors in, separated "
                                       f"waveforms [B,N+1,65280] out to pinned host memory; {EVAL_DEPTH} steps in flight on {EVAL_DEPTH} streams",
                 waveforms_stay_on_device=dict(value=per(ms_e2e_dev), ms_per_step=ms_e2e_dev / args.steps, d2h_bytes_per_step=0,
                                               note="same call with wav_host = NULL: what a device-side evaluator consumes")),
        metrics=dict(ms_per_step=ms_met, clips_per_s=B * world / (ms_met * 1e-3), mean_sdr_db=mean_sdr,
                     clips_per_s_separation_plus_metrics=B * world / ((ms_total / args.steps + ms_met) * 1e-3),
                     evaluator="avsep_bss_eval: hand-written fp64 kernels (direct lag correlations + block-Toeplitz Levinson + quadratic-form energies), all clips of the batch in one call; no cuFFT / cuSOLVER"),
        roofline=roofline, kernels=kernels,
        collective=("all_reduce([sdr, sir, sar, si_sdr sums, count]) over NCCL, async, every step" if world > 1 else None))


def torch_gpu_baseline(dev):
    """The reference's own call sequence with stock PyTorch kernels on this GPU ("the real bar", since the
    reference ships no Blackwell kernel): oracle/torch_ref_path.py -- the code the CPU arm times -- with every tensor on cuda."""
    from oracle import torch_ref_path as TP
    out = {}
    try:
        for name, (B, N, fn, steps) in {"train_C2": (B_PER_GPU, N_SRC, "train", 5), "eval_C3": (128, 3, "eval", 2)}.items():
            waves = make_waves(B, N, 1235).permute(1, 0, 2).contiguous().to(dev)      # [B, N+1, L]
            g = torch.Generator().manual_seed(1235)
            feat = torch.randn(B, K_CH, 256, 256, generator=g).to(dev)
            vecs = torch.sigmoid(torch.randn(B, N, K_CH, generator=g)).to(dev)
            scale = torch.ones(K_CH, device=dev, requires_grad=True)
            bias = torch.zeros(1, device=dev, requires_grad=True)

            def call():
                if fn == "train":
                    f, v = feat.clone().requires_grad_(True), vecs.clone().requires_grad_(True)
                    scale.grad = None
                    bias.grad = None
                    return TP.train_step(waves, f, v, scale, bias)
                return TP.eval_step(waves, feat, vecs, scale.detach(), bias.detach())

            call()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            out[name] = dict(batch=B, num_mix=N, ms_per_step=ms, value=B / ms * 1e3, unit=UNIT, steps=steps)
            del waves, feat, vecs
            torch.cuda.empty_cache()
        out["what"] = (f"stock PyTorch {torch.__version__} on this GPU, the reference's call sequence (torch.stft / F.grid_sample with "
                       "host-built grids / bmm / sigmoid / F.binary_cross_entropy / autograd / torch.istft in fp64); no SDR metrics")
    except Exception as e:                                                        # pragma: no cover
        out["error"] = repr(e)
    return out


def main_ours(args, rank, world, local):
    import torch.distributed as dist
    import avsep_b200 as av
    from avsep_b200 import parallel

    torch.cuda.set_device(local)
    numa_node = parallel.bind_to_gpu_numa(local) if world > 1 else None   # node-local pinned buffers for the H2D copies
    av.device_check()
    dev = torch.device("cuda", local)
    pk = peaks()
    main_stream = torch.cuda.Stream()                              # graphs cannot be captured on the legacy default stream
    torch.cuda.set_stream(main_stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    tr, ev = None, None
    if args.workload in ("both", "train"):
        tr = train_section(av, parallel, dist, args, rank, world, local, dev, pk, sampler)
    torch.cuda.empty_cache()
    if args.workload in ("both", "eval"):
        ev = eval_section(av, parallel, dist, args, rank, world, dev, pk)
        torch.cuda.empty_cache()
    clocks = tr["clocks"] if tr else (sampler.stop() if rank == 0 else None)

    # ---- baselines on the box (rank 0, N = 1 only): stock torch on the GPU, the reference torch path on the host cores ----
    gpu_baseline, cpu_baseline = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_baseline = torch_gpu_baseline(dev)
        try:
            v, times_cpu = run_cpu_baseline(16, 5, budget_s=25.0)
            cpu_baseline = dict(value=v, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                sample=f"16 mixtures per step, best of {len(times_cpu)} steps after 1 warm-up, same C2 train-step workload "
                                       f"(oracle/torch_ref_path.py: the reference's torch call sequence; torch.stft for the absent librosa)")
        except Exception as e:                                                    # pragma: no cover
            cpu_baseline = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {e}")

    if rank == 0:
        if tr:
            B, N, K = B_PER_GPU, N_SRC, K_CH
            line = dict(
                metric=METRIC, value=tr["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=tr["ms_total"] / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="fp32 (DFT GEMM: bf16x3 split operands, fp32 TMEM accumulate)", data="synthetic",
                config=dict(workload=("C2 MUSIC N=2 train step: STFT (2 source transforms, mixture spectrum = X0 + X1) + prepare + synth K=32 "
                                      "+ sigmoid + weighted BCE + backward" if tr["src_only"] else
                                      "C2 MUSIC N=2 train step: STFT (3 transforms: explicit mixture waveform) + prepare + synth K=32 + sigmoid "
                                      "+ weighted BCE + backward"),
                            batch_per_gpu=B, global_batch=B * world, num_mix=N, K=K, samples=L, stft="1022/256", grid="256x256",
                            parallelism=f"batch-sharded x{world}, no data-path collective" +
                                        ("; one async NCCL all-reduce of [loss, count] per step inside the timed loops" if world > 1 else ""),
                            numa_node_rank0=numa_node,
                            launch=("plain launches" if args.no_graph else "CUDA graph replay of the step's kernels (see gpu_launches / steps)"),
                            mixture=("explicit mixture waveform, 3 transforms" if not tr["src_only"] else
                                     "formed on the GPU from the 2 device-resident source waveforms (dataset/music.py:127): mixture "
                                     "spectrum = X0 + X1, 2 transforms executed"),
                            l2="inputs larger than L2: 1.3 GB of feat_sound/dfeat streamed per step vs 126 MB L2"),
                e2e=tr["e2e"], value_explicit_mixture=tr["explicit"], value_pipelined=tr["pipelined"], gpu_launches=tr["launches"], clocks=clocks, loss=tr["loss"],
                loss_global_mean=tr["loss_global"], loss_e2e=tr["loss_e2e"], collective=tr["collective"], roofline=tr["roofline"],
                kernels=tr["kernels"], eval=ev, cpu_baseline=cpu_baseline, torch_gpu_baseline=gpu_baseline)
        else:
            line = dict(metric=METRIC, value=ev["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                        ms_per_step=ev["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                        dtype="fp32 (DFT GEMM: bf16x3 split operands, fp32 TMEM accumulate)", data="synthetic", config=ev["config"],
                        e2e=ev["e2e"], gpu_launches=ev["gpu_launches"], clocks=clocks, roofline=ev["roofline"], kernels=ev["kernels"],
                        metrics=ev["metrics"], cpu_baseline=cpu_baseline, torch_gpu_baseline=gpu_baseline)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "train", "eval"],
                    help="both (default): the C2 train-step line with the C3 eval block inside; train / eval: one of them only")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core and stock-torch-GPU baselines")
    ap.add_argument("--no-graph", action="store_true", help="plain kernel launches instead of CUDA-graph replay of the step")
    ap.add_argument("--explicit-mixture", action="store_true",
                    help="train workload: hand the step the mixture waveform as a third signal instead of letting it form "
                         "the mixture from the sources on the GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep stdout to the one JSON line
        from avsep_b200 import parallel
        parallel.init_from_env("nccl")
    main_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
