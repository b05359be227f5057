#!/usr/bin/env python
"""Benchmark of the mix-and-separate spectrogram hot path (BASELINE.json: separated mixtures/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|eval]

Workload (config C2, BASELINE.json configs[1]): MUSIC N=2 training step, batch 80 per GPU:
STFT of the (N+1) signals of every mixture (65535 samples @ 11025 Hz, n_fft 1022 / hop 256) ->
prepare (log-frequency warp 512->256, weights, ground-truth masks, log magnitude) -> K=32
synthesizer masks + sigmoid + weighted BCE -> full backward (dfeat_sound, dvecs, dscale, dbias).
Synthetic MUSIC-shaped inputs (harmonic tones); feat_sound ~ N(0,1) stands in
for the U-Net output and is device-resident, exactly as in the reference.

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = the same step through
the host-buffer C-ABI entry point (pinned waveforms + K-vectors H2D, loss D2H inside the timed region).
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

# algorithmic bytes / flops per mixture at N=2 (each boundary tensor touched once)
ALG = {
    "dft_gemm_fwd": dict(bound="tensor", flops=1.607e9),                       # 3 signals x 0.536 GFLOP
    "stft_fwd": dict(bound="tensor", flops=1.607e9),                           # staging kernel + DFT GEMM together
    "wave_split": dict(bound="hbm", bytes=786420 + 3 * 66560 * 4),             # fp32 in + bf16 hi/lo out
    # fold staging: fp32 waves in + folded E / O operand out (3 signals x 256 frames x 512 x 4 planes x 2 B)
    "wave_fold": dict(bound="hbm", bytes=786420 + 3 * 256 * 512 * 4 * 2),
    "prepare": dict(bound="hbm", bytes=1572864 + 786432 + 1048576),            # warp in + out + derived out
    # fused forward + loss + backward: feat_sound in (once), dfeat out, masks out, gt + weight in, K-vectors
    "synth_train": dict(bound="hbm", bytes=8388864 + 8388608 + 524288 + 524288 + 262144),
    # eval-only kernels: un-warp in + (mix re, im) in -> bf16 hi/lo operand out (3 signals x 256 x 1024 x 2 x 2 B)
    "spec_prep": dict(bound="hbm", bytes=524288 + 1048576 + 3 * 256 * 1024 * 4),
    "dft_gemm_inv": dict(bound="tensor", flops=1.607e9),                       # 3 signals x 0.536 GFLOP
    # the unfused pair (reported for comparison only): SURVEY.md section 8(d) staged figures
    "synth_fwd_loss": dict(bound="hbm", bytes=8388864 + 524288 + 1310720),
    "synth_bwd": dict(bound="hbm", bytes=9699328 + 8388608),
}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel, from the committed `ncu --set full` summary
    (profiles/r1_ncu_full_summary.txt, produced by tools/ncu_summary.py); None if the kernel is not in it."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_summary.txt")
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
def main_ours(args, rank, world, local):
    import torch.distributed as dist
    import avsep_b200 as av
    from avsep_b200 import parallel

    torch.cuda.set_device(local)
    numa_node = parallel.bind_to_gpu_numa(local) if world > 1 else None   # node-local pinned buffers for the H2D copies
    av.device_check()
    dev = torch.device("cuda", local)
    B, N, K = B_PER_GPU, N_SRC, K_CH
    pk = peaks()

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
    loss = torch.zeros(1, device=dev)
    dvecs_host = torch.empty((B, N, K)).pin_memory()
    step = av.Step(B, N, K, L)
    step.set_graph(not args.no_graph)                              # train step replayed as a CUDA graph (side streams only)
    eval_mode = args.workload == "eval"
    if not eval_mode and not args.explicit_mixture:
        step.set_mix_on_device(True)
    main_stream = torch.cuda.Stream()                              # graphs cannot be captured on the legacy default stream
    torch.cuda.set_stream(main_stream)
    wav_dev = torch.empty((B, N + 1, 65280), device=dev) if eval_mode else None
    wav_host = torch.empty((B, N + 1, 65280)).pin_memory() if eval_mode else None

    # Train step: the step is handed the SOURCE waveforms only and forms the mixture on the GPU, as the e2e arm does
    # (dataset/music.py:127 defines the mixture as the sum of its sources); --explicit-mixture hands over the
    # mixture waveform as a third signal instead.
    src_only = not eval_mode and not args.explicit_mixture
    sources_dev = waves_dev[1:]                                    # [N, B, L] view

    def step_dev():
        if eval_mode:
            step.eval_dev(waves_dev, vecs_dev, feat, scale, bias, wav_dev)
        elif src_only:
            step.train_dev(sources_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)
        else:
            step.train_dev(waves_dev, vecs_dev, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)

    def step_host():
        if eval_mode:
            step.eval_host(waves_host, vecs_host, feat, scale, bias, wav_host)
            return 0.0
        return step.train_host(waves_host, vecs_host, feat, scale, bias, dfeat, dvecs_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ------------------------------------------------------------
    # The clock sampler (nvidia-smi -lms 200) runs from before the warm-up to after the e2e loop so its samples
    # bracket both timed regions.  Its start-up (NVML attach) perturbs kernel launches for a few hundred ms, so the
    # warm-up keeps the GPU loaded until the sampler has produced its first samples (bounded at 4 s).
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
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
    barrier()
    launches0 = av.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    launches = av.launch_count() - launches0
    loss_val = float(loss)
    value = B * world * args.steps / (ms_total * 1e-3)

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
    # region.  Two Step objects alternate on two streams (a depth-2 software pipeline, like a prefetching
    # training loop), so the copy of step i+1 overlaps the kernels of step i; the loss of step i is read on
    # the host before its Step object is reused for step i+2.
    if eval_mode:
        def run_e2e(k, mix_on_device=False):
            for _ in range(k):
                step_host()
            return 0.0
        pipe_depth = 1
    else:
        step_b = av.Step(B, N, K, L)
        step_b.set_graph(not args.no_graph)
        pipes = [dict(step=step, stream=torch.cuda.Stream(), dfeat=dfeat, loss=torch.zeros(1).pin_memory(),
                      dvecs=torch.empty((B, N, K)).pin_memory()),
                 dict(step=step_b, stream=torch.cuda.Stream(), dfeat=torch.empty_like(feat), loss=torch.zeros(1).pin_memory(),
                      dvecs=torch.empty((B, N, K)).pin_memory())]
        pipe_depth = 2
        sources_host = waves_host[1:]                              # [N, B, L] view of the same pinned buffer

        def run_e2e(k, mix_on_device):
            for p in pipes:
                p["step"].set_mix_on_device(mix_on_device)
            src = sources_host if mix_on_device else waves_host
            for i in range(k):
                p = pipes[i % 2]
                p["stream"].synchronize()                          # step i-2 done: its loss is on the host now
                _ = float(p["loss"][0])
                with torch.cuda.stream(p["stream"]):
                    p["step"].train_host_async(src, vecs_host, feat, scale, bias, p["dfeat"], p["loss"], p["dvecs"])
            for p in pipes:
                p["stream"].synchronize()
            return float(pipes[(k - 1) % 2]["loss"][0])

    def timed_e2e(mix_on_device):
        torch.cuda.synchronize()
        run_e2e(4, mix_on_device)
        barrier()
        t0 = time.perf_counter()
        lh = run_e2e(args.steps, mix_on_device)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return parallel.max_over_ranks(wall, device=dev), lh

    # primary: the host supplies the N source waveforms, the mixture (their sum, dataset/music.py:127) is formed on
    # the GPU inside the staging kernel; secondary: the host also supplies the mixture waveform (N+1 signals H2D)
    ms_e2e, loss_host = timed_e2e(True)
    ms_e2e_full, loss_host_full = timed_e2e(False) if not eval_mode else (ms_e2e, loss_host)
    if not eval_mode:
        for p in pipes:
            p["step"].set_mix_on_device(False)
        step.set_mix_on_device(src_only)             # pipes[0] is the value arm's Step: back to the input layout step_dev() passes
        # same inputs, same kernels through the host-buffer entry point: the losses must agree (a gross mismatch = a bug;
        # the exact values are reported in the JSON line)
        assert abs(loss_host - loss_val) <= 1e-3 * max(abs(loss_val), 1.0), (loss_host, loss_val)
        # explicit mixture waveform vs |X0 + X1|: equal to fp32 rounding of the magnitudes (a few binary targets may flip)
        assert abs(loss_host_full - loss_val) <= 1e-2 * max(abs(loss_val), 1.0), (loss_host_full, loss_val)
    for _ in range(40):                                            # keep the GPU loaded until a few more samples land
        step_dev()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)
    e2e_full_value = B * world * args.steps / (ms_e2e_full * 1e-3)
    h2d_full = waves_host.numel() * 4 + vecs_host.numel() * 4
    h2d = (h2d_full if eval_mode else waves_host[1:].numel() * 4 + vecs_host.numel() * 4)
    d2h = (wav_host.numel() * 4) if eval_mode else (4 + dvecs_host.numel() * 4)
    if world > 1:
        gl = parallel.global_mean_loss(torch.tensor(loss_val, device=dev), B)   # the path's only collective: 2 scalars
        loss_val = float(gl)

    # ---- per-kernel roofline (rank 0), each kernel timed alone with CUDA events ---------------------------
    kernels = []
    roofline = None
    if rank == 0 and eval_mode:
        lib, stream, chk = av._lib.lib, torch.cuda.current_stream().cuda_stream, av._lib.check
        plan = lib.avsep_step_plan(step._h)
        S = (N + 1) * B
        mag = torch.empty((S, 512, T), device=dev)
        re = torch.empty((B, 512, T), device=dev)
        im = torch.empty((B, 512, T), device=dev)
        pred = step.buffer("pred_masks")
        t_stft = time_calls(lambda: chk(lib.avsep_stft(plan, waves_dev.data_ptr(), S, mag.data_ptr(), None, None, None, stream)), 10)
        t_gemm = time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan, 0, S, mag.data_ptr(), stream)), 10)
        step.eval_dev(waves_dev, vecs_dev, feat, scale, bias, wav_dev)
        re.copy_(step.buffer("mix_re")); im.copy_(step.buffer("mix_im"))
        t_sep = time_calls(lambda: chk(lib.avsep_separate(plan, pred.data_ptr(), N, B, None, None, re.data_ptr(), im.data_ptr(), 1, 1,
                                                          0.5, 1, wav_dev.data_ptr(), stream)), 10)
        t_inv = time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan, 1, S, wav_dev.data_ptr(), stream)), 10)
        times = {"dft_gemm_fwd": t_gemm, "wave_split": max(t_stft - t_gemm, 1e-6), "spec_prep": max(t_sep - t_inv, 1e-6),
                 "dft_gemm_inv": t_inv}
        if t_stft - t_gemm < 0.03 * t_stft:      # fused staging (dft_fold3): the forward transform is ONE kernel
            times["dft_gemm_fwd"] = t_stft
            del times["wave_split"]
        for name, ms in times.items():
            a = ALG[name]
            if a["bound"] == "tensor":
                ach = a["flops"] * B / (ms * 1e-3) / 1e12
                kernels.append(dict(kernel=name, ms=ms, bound="tensor", achieved=ach, peak=pk["tensor"], unit="TFLOP/s",
                                    frac=ach / pk["tensor"], executed_tflops=3 * ach, tensor_pipe_frac=3 * ach / pk["tensor"]))
            else:
                ach = a["bytes"] * B / (ms * 1e-3) / 1e9
                kernels.append(dict(kernel=name, ms=ms, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"]))
        top = max(kernels, key=lambda k: k["ms"])
        roofline = dict(kernel=top["kernel"], bound=top["bound"], achieved=top["achieved"], peak=top["peak"], unit=top["unit"],
                        frac=top["frac"], traffic=ncu_traffic({"dft_gemm_inv": "dft_gemm2_kernel", "spec_prep": "spec_prep_kernel",
                                                               "dft_gemm_fwd": "dft_fold3_fwd_kernel"}.get(top["kernel"], top["kernel"])),
                        traffic_unit="bytes per launch (dram read + write), ncu --set full: profiles/r1_ncu_full_summary.txt",
                        peak_source=pk["source"])
    if rank == 0 and not eval_mode:
        lib, stream = av._lib.lib, torch.cuda.current_stream().cuda_stream
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
        chk = av._lib.check
        if src_only:                                 # one kernel: two source transforms, mixture spectrum = their sum
            t_stft = time_calls(lambda: chk(lib.avsep_stft_mix(plan, sources_dev.data_ptr(), B, N, mag.data_ptr(), None, stream)), 10)
            t_gemm = t_stft
        else:
            t_stft = time_calls(lambda: chk(lib.avsep_stft(plan, waves_dev.data_ptr(), S, mag.data_ptr(), None, None, None, stream)), 10)
            t_gemm = time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan, 0, S, mag.data_ptr(), stream)), 10)
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
        times = {"dft_gemm_fwd": t_gemm, "wave_fold": max(t_stft - t_gemm, 1e-6), "prepare": t_prep, "synth_train": t_train}
        if t_stft - t_gemm < 0.03 * t_stft:      # fused staging (dft_fold3): the forward transform is ONE kernel
            times["dft_gemm_fwd"] = t_stft
            del times["wave_fold"]
        unfused = {"synth_fwd_loss": t_fwd, "synth_bwd": t_bwd}

        tot = sum(times.values())
        for name, ms in times.items():
            a = ALG[name]
            if a["bound"] == "tensor":
                ach = a["flops"] * B / (ms * 1e-3) / 1e12
                kernels.append(dict(kernel=name, ms=ms, share=ms / tot, bound="tensor", achieved=ach, peak=pk["tensor"],
                                    unit="TFLOP/s", frac=ach / pk["tensor"], executed_tflops=(1.0 if src_only else 1.5) * ach,
                                    tensor_pipe_frac=(1.0 if src_only else 1.5) * ach / pk["tensor"],
                                    note="algorithmic = fp32-equivalent full DFT (0.536 GFLOP/signal); executed = bf16x3 "
                                         "split on the folded (half-size) contraction = 1.5x algorithmic; with fused "
                                         "staging (dft_fold3) this one kernel also folds / splits the waveform; when the "
                                         "step forms the mixture from its 2 sources (default) only 2 of the 3 transforms "
                                         "are executed (mixture spectrum = X0 + X1), i.e. executed = 1.0x algorithmic"))
            else:
                ach = a["bytes"] * B / (ms * 1e-3) / 1e9
                kernels.append(dict(kernel=name, ms=ms, share=ms / tot, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s",
                                    frac=ach / pk["hbm"]))
        for name, ms in unfused.items():
            ach = ALG[name]["bytes"] * B / (ms * 1e-3) / 1e9
            kernels.append(dict(kernel=name, ms=ms, share=None, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s",
                                frac=ach / pk["hbm"], note="not part of the step: the unfused forward / backward pair, for comparison"))
        top = max([k for k in kernels if k["share"] is not None], key=lambda k: k["ms"])
        roofline = dict(kernel=top["kernel"], bound=top["bound"], achieved=top["achieved"], peak=top["peak"], unit=top["unit"],
                        frac=top["frac"], traffic=ncu_traffic({"synth_train": "synth_train_kernel", "dft_gemm_fwd": "dft_fold3m_fwd_kernel",
                                                               "prepare": "prepare_vec_kernel"}.get(top["kernel"], top["kernel"])),
                        traffic_unit="bytes per launch (dram read + write), ncu --set full: profiles/r1_ncu_full_summary.txt",
                        algorithmic_bytes_per_launch=(ALG[top["kernel"]].get("bytes", 0) * B or None),
                        peak_source=pk["source"],
                        note=("algorithmic fp32-equivalent FLOPs (0.536 GFLOP/signal)" if top["bound"] == "tensor" else
                              "algorithmic bytes per SURVEY.md section 8(d) (feat_sound counted once: fused fwd+bwd)"))

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample ------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, times_cpu = run_cpu_baseline(16, 5, budget_s=25.0)
            cpu_baseline = dict(value=v, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                sample=f"16 mixtures per step, best of {len(times_cpu)} steps after 1 warm-up, same C2 train-step workload "
                                       f"(oracle/torch_ref_path.py: the reference's torch call sequence; torch.stft for the absent librosa)")
        except Exception as e:                                                    # pragma: no cover
            cpu_baseline = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {e}")

    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="fp32 (DFT GEMM: bf16x3 split operands, fp32 TMEM accumulate)", data="synthetic",
            config=dict(workload=("C3-style eval step (STFT+prepare+synth+sigmoid+BCE+unwarp+mask+iSTFT)" if eval_mode else
                                  "C2 MUSIC N=2 train step: STFT(3 signals)+prepare+synth K=32+sigmoid+weighted BCE+backward"),
                        batch_per_gpu=B, global_batch=B * world, num_mix=N, K=K, samples=L, stft="1022/256", grid="256x256",
                        parallelism=f"batch-sharded x{world}, no data-path collective", numa_node_rank0=numa_node,
                        launch=("plain launches" if args.no_graph or eval_mode else "CUDA graph replay of the step's kernels (see gpu_launches / steps)"),
                        mixture=("n/a" if eval_mode else
                                 "explicit mixture waveform, 3 transforms" if args.explicit_mixture else
                                 "formed on the GPU from the 2 device-resident source waveforms (dataset/music.py:127): mixture "
                                 "spectrum = X0 + X1, 2 transforms executed"),
                        l2="inputs larger than L2: 1.3 GB of feat_sound/dfeat streamed per step vs 126 MB L2"),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps,
                     pipeline_depth=pipe_depth, timer="host wall clock around the K steps, synchronised on both sides",
                     api="avsep_step_train_host_async (C-ABI): pinned host SOURCE waveforms [N,B,L] + K-vectors in, mixture summed "
                         "on the GPU (avsep_step_set_mix_on_device), loss + dvecs out; 2 steps in flight on 2 streams",
                     with_mixture_h2d=dict(value=e2e_full_value, ms_per_step=ms_e2e_full / args.steps, h2d_bytes_per_step=h2d_full,
                                           note="same call with the mixture waveform also supplied by the host, [N+1,B,L]")),
            value_explicit_mixture=explicit, gpu_launches=int(launches), clocks=clocks, loss=loss_val,
            loss_e2e=(None if eval_mode else dict(sources_only=loss_host, explicit_mixture=loss_host_full)), roofline=roofline, kernels=kernels, cpu_baseline=cpu_baseline)
        print(json.dumps(line), flush=True)
    step.close()
    if not eval_mode:
        step_b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "eval"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
