#!/usr/bin/env python
"""Stock-PyTorch GPU baseline for the hot path ("the real bar").

Runs the reference's own sequence of torch calls (oracle/torch_ref_path.py -- the same code the CPU arm
times) with every tensor on cuda:0, i.e. what the unmodified reference would do on a B200 if its data
loader handed it waveforms: torch.stft (cuFFT), F.grid_sample with host-built grids, bmm, sigmoid,
F.binary_cross_entropy, autograd backward, torch.istft in fp64.  Same synthetic workload as bench.py.
Prints one JSON line; not part of the product path and not a bench.py arm.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402  (workload generators only)
from oracle import torch_ref_path as ref          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=80)
    ap.add_argument("--mix", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, N, K = args.batch, args.mix, 32
    waves = bench.make_waves(B, N, 1235).permute(1, 0, 2).contiguous().to(dev)   # [B, N+1, L], index 0 = mixture
    g = torch.Generator().manual_seed(1235)
    feat = torch.randn(B, K, 256, 256, generator=g).to(dev).requires_grad_(True)
    vecs = torch.sigmoid(torch.randn(B, N, K, generator=g)).to(dev).requires_grad_(True)
    scale = torch.ones(K, device=dev, requires_grad=True)
    bias = torch.zeros(1, device=dev, requires_grad=True)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    def train():
        feat.grad = None
        vecs.grad = None
        scale.grad = None
        bias.grad = None
        return ref.train_step(waves, feat, vecs, scale, bias)

    def evaluate():
        return ref.eval_step(waves, feat.detach(), vecs.detach(), scale.detach(), bias.detach())

    loss = train()
    t_train = timed(train)
    t_eval = timed(evaluate)
    print(json.dumps({
        "what": "stock PyTorch %s on cuda:0 (%s), reference call sequence" % (torch.__version__, torch.cuda.get_device_name(0)),
        "batch": B, "num_mix": N, "loss": loss,
        "train_ms_per_step": t_train, "train_mixtures_per_s": B / t_train * 1e3,
        "eval_ms_per_step": t_eval, "eval_mixtures_per_s": B / t_eval * 1e3,
        "steps": args.steps, "warmup": args.warmup,
        "note": "train_step includes a float(loss) D2H sync per step, as the reference's err.item() does; grids are rebuilt on the host and copied every call, as main.py:62-63 does"}))


if __name__ == "__main__":
    main()
