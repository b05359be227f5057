#!/usr/bin/env python
"""HBM ceilings by access mix on this GPU (CUDA events, best of 10, 1 GiB buffers): copy (read + write), fill (write only),
reduce (read only).  Context for the roofline fractions of write-dominated kernels (spec_prep, the operand-staging kernels)."""
import json

import torch

dev = torch.device("cuda", 0)
n = 1 << 28                                                       # 1 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
a.normal_()


def best(fn, reps=10):
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


out = {}
out["copy_gbs"] = 2 * 4 * n / best(lambda: b.copy_(a)) / 1e6
out["fill_gbs"] = 4 * n / best(lambda: b.fill_(1.0)) / 1e6
out["memset_gbs"] = 4 * n / best(lambda: b.zero_()) / 1e6
out["read_sum_gbs"] = 4 * n / best(lambda: a.sum()) / 1e6
print(json.dumps(out))
