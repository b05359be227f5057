#!/usr/bin/env python
"""SM-partitioned front / back pipeline probe: does the tensor-bound STFT (K1) of batch i + 1 overlap the
HBM-bound prepare + synthesizer kernels (K2 + K3) of batch i when the two run on disjoint SM partitions (CUDA green contexts)?

For each front partition size it reports, per batch of B mixtures (C2 train step, N = 2, K = 32):
  k1_alone      K1 on the front partition, nothing else running
  k23_alone     K2 + K3 + finalize on the back partition, nothing else running
  pipelined     the dependent two-stream pipeline (K1 of batch i + 1 beside K2 + K3 of batch i, two mag buffers), whole-loop time / n
  serial_full   the same kernels back to back on one ordinary stream with all SMs (what the step does today)
Rank-0 JSON on stdout.  Bits do not depend on the partition (torch.equal check against the serial result)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av  # noqa: E402,F401
import bench  # noqa: E402
from avsep_b200 import _lib, ops  # noqa: E402
from cuda.bindings import driver as drv  # noqa: E402

lib, check, ptr_array = _lib.lib, _lib.check, _lib.ptr_array
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B, N, K, L = int(os.environ.get("PROBE_B", "80")), 2, 32, 65535
ITERS = int(os.environ.get("PROBE_ITERS", "30"))


def ck(r):
    if r[0] != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"driver call failed: {r[0]}")
    return r[1] if len(r) == 2 else r[1:]


def green_streams(front_sms):
    """(front stream, front SMs, back stream, back SMs): two green contexts that partition the device's SMs"""
    torch.zeros(1, device=dev)                                    # primary context
    cu_dev = ck(drv.cuDeviceGet(0))
    res = ck(drv.cuDeviceGetDevResource(cu_dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    groups, _, rem = ck(drv.cuDevSmResourceSplitByCount(1, res, 0, front_sms))
    out = []
    for r in (groups[0], rem):
        desc = ck(drv.cuDevResourceGenerateDesc([r], 1))
        g = ck(drv.cuGreenCtxCreate(desc, cu_dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = ck(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out += [torch.cuda.get_stream_from_external(int(s), dev), int(r.sm.smCount)]
    return out


waves = bench.make_waves(B, N, 1234).to(dev)
src = waves[1:].contiguous()                                      # [N, B, L]
g = torch.Generator().manual_seed(7)
feat = torch.randn(B, K, 256, 256, generator=g).to(dev)
vecs = torch.sigmoid(torch.randn(B, N, K, generator=g)).to(dev)
scale, bias = torch.ones(K, device=dev), torch.zeros(1, device=dev)
plan = ops.Plan((N + 1) * B, L)
T = plan.frames
mags = [torch.empty((N + 1) * B, 512, T, device=dev) for _ in range(2)]
img = lambda n=1: torch.empty(n * B, 256, T, device=dev)
mix_w, log_mix, weights, mags_w, gt, pred = img(), img(), img(), img(N), img(N), img(N)
dfeat = torch.empty_like(feat)
dvecs, dscale, dbias, loss = torch.empty_like(vecs), torch.empty(K, device=dev), torch.empty(1, device=dev), torch.zeros(1, device=dev)
scratch = torch.empty(int(lib.avsep_synth_scratch_bytes(B, N, K, 256 * T)), dtype=torch.uint8, device=dev)
sig, im = 512 * T * 4, 256 * T * 4


def k1(slot, stream):
    check(lib.avsep_stft_mix(plan.handle, src.data_ptr(), B, N, mags[slot].data_ptr(), None, stream.cuda_stream), "stft_mix")


def k23(slot, stream):
    m = mags[slot].data_ptr()
    s = stream.cuda_stream
    check(lib.avsep_prepare(m, ptr_array([m + (1 + n) * B * sig for n in range(N)]), N, B, T, 1, 1, 1, mix_w.data_ptr(),
                            ptr_array([mags_w.data_ptr() + n * B * im for n in range(N)]), log_mix.data_ptr(), weights.data_ptr(),
                            ptr_array([gt.data_ptr() + n * B * im for n in range(N)]), s), "prepare")
    check(lib.avsep_synth_train(feat.data_ptr(), vecs.data_ptr(), scale.data_ptr(), bias.data_ptr(), B, N, K, 256 * T, 1, 3,
                                gt.data_ptr(), weights.data_ptr(), None, pred.data_ptr(), loss.data_ptr(), dfeat.data_ptr(),
                                dvecs.data_ptr(), dscale.data_ptr(), dbias.data_ptr(), scratch.data_ptr(), s), "synth_train")


def timed(stream, fn, iters=ITERS, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def pipelined(front, back, iters=ITERS, warm=5):
    """dependent pipeline over iters batches: front K1(i) -> mags[i % 2]; back K2 + K3(i) after K1(i); K1(i + 2) after the back of i"""
    def run(n):
        fdone = [torch.cuda.Event() for _ in range(n)]
        bdone = [torch.cuda.Event() for _ in range(n)]
        for i in range(n):
            if i >= 2:
                front.wait_event(bdone[i - 2])
            k1(i % 2, front)
            fdone[i].record(front)
            back.wait_event(fdone[i])
            k23(i % 2, back)
            bdone[i].record(back)
    run(warm)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    front.wait_stream(main)
    back.wait_stream(main)
    run(iters)
    main.wait_stream(front)
    main.wait_stream(back)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = dict(B=B, N=N, iters=ITERS, splits=[])
plain = torch.cuda.Stream()
plan.set_sm_budget(0)
out["k1_full"] = timed(plain, lambda: k1(0, plain))
out["k23_full"] = timed(plain, lambda: k23(0, plain))
out["serial_full"] = timed(plain, lambda: (k1(0, plain), k23(0, plain)))
torch.cuda.synchronize()
ref = (loss.clone(), dfeat.clone(), pred.clone())
# same two-stream dependent pipeline WITHOUT partitioning (ordinary streams, all SMs): what the block scheduler does on its own
s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
out["pipelined_unpartitioned"] = pipelined(s_a, s_b)
for budget in [int(x) for x in os.environ.get("PROBE_BUDGETS", "48,64,80").split(",")]:
    plan.set_sm_budget(budget)
    out[f"pipelined_unpartitioned_k1_budget_{budget}"] = pipelined(s_a, s_b)
for front_sms in [int(x) for x in os.environ.get("PROBE_SPLITS", "40,48,56,64,72").split(",")]:
    try:
        front, nf, back, nb = green_streams(front_sms)
    except Exception as e:  # noqa: BLE001
        out["splits"].append(dict(front_sms=front_sms, error=str(e)))
        continue
    plan.set_sm_budget(nf)
    r = dict(front_sms=nf, back_sms=nb)
    r["k1_alone"] = timed(front, lambda: k1(0, front))
    r["k23_alone"] = timed(back, lambda: k23(0, back))
    r["pipelined"] = pipelined(front, back)
    torch.cuda.synchronize()
    r["bits_equal_serial"] = bool(torch.equal(loss, ref[0]) and torch.equal(dfeat, ref[1]) and torch.equal(pred, ref[2]))
    r["mixtures_per_s"] = B / r["pipelined"] * 1e3
    out["splits"].append(r)
plan.set_sm_budget(0)
print(json.dumps(out))
