"""Diagnostic: per-step device times of the C2 train step over a long run, with / without the nvidia-smi sampler."""
import os, sys, time, statistics, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avsep_b200 as av
import bench

B, N, K, L = 80, 2, 32, 65535
dev = torch.device("cuda", 0)
waves = bench.make_waves(B, N, 1234).to(dev)
g = torch.Generator(device=dev).manual_seed(99)
feat = torch.randn((B, K, 256, 256), device=dev, generator=g)
vecs = torch.sigmoid(torch.randn((B, N, K), device=dev, generator=g))
scale = torch.ones(K, device=dev); bias = torch.zeros(1, device=dev)
dfeat = torch.empty_like(feat); dvecs = torch.empty((B, N, K), device=dev)
dscale = torch.empty(K, device=dev); dbias = torch.empty(1, device=dev); loss = torch.zeros(1, device=dev)
step = av.Step(B, N, K, L)

def run(n, tag, sampler=False):
    proc = None
    if sampler:
        proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        time.sleep(0.3)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        step.train_dev(waves, vecs, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    out = ""
    if proc:
        proc.terminate()
        out = proc.communicate()[0].strip().splitlines()
    q = lambda a: f"min {min(a):.3f} med {statistics.median(a):.3f} mean {statistics.mean(a):.3f} max {max(a):.3f}"
    print(f"{tag}: n={n}  all[{q(ms)}]  first50[{q(ms[:50])}]  last50[{q(ms[-50:])}]")
    if out:
        print("   smi:", " | ".join(out[:3]), "...", " | ".join(out[-3:]))
    big = [(i, round(m, 2)) for i, m in enumerate(ms) if m > 1.5 * statistics.median(ms)]
    print("   outliers(>1.5x median):", big[:20], "count", len(big))

run(20, "cold20")
run(200, "burst200")
time.sleep(2)
run(3000, "long3000")
run(3000, "long3000+smi", sampler=True)
time.sleep(3)
run(20, "after-idle20")
