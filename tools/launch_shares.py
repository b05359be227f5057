#!/usr/bin/env python
"""Per-kernel launch counts / average durations from an `ncu --metrics gpu__time_duration.sum --csv` launch list, and each
train-step kernel's share of the step (runs here, no GPU).  usage: launch_shares.py gpurun_out/r2_ncu_launches.csv"""
import csv
import sys
from collections import defaultdict

STEP = ("dft_fold3m_fwd_kernel<2, 2>", "prepare_bulk_kernel<2>", "synth_train_kernel", "synth_finalize_kernel")
EVAL = ("dft_fold3m_fwd_kernel<2, 3>", "prepare_bulk_kernel<3>", "synth_fwd_staged_kernel<3", "spec_prep_tma_kernel", "idft_fold2_kernel")

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(list)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    agg[r[ki]].append(v)


def avg(sub):
    for k, v in agg.items():
        if sub in k:
            return sum(v) / len(v)
    return None


print(f"# {sys.argv[1]}: per-launch times under ncu are cold-cache and serialised -- compare SHARES, not absolutes")
for group, names in (("train step (C2, B = 80, N = 2)", STEP), ("eval step (C3, B = 128, N = 3)", EVAL)):
    t = {n: avg(n) for n in names}
    tot = sum(v for v in t.values() if v)
    print(f"# {group}: sum of the step's kernels {tot:.1f} us")
    for n, v in t.items():
        if v:
            print(f"#   {n:28s} {v:9.2f} us   share {v / tot:.3f}")
print()
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:110]:110s} n={len(v):4d} avg={sum(v) / len(v):10.2f} us")
