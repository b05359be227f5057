"""bss_metrics.bss_eval_batch (the batched device evaluator) against the CPU
restatement of mir_eval / pb_bss_eval in oracle/bss_eval.py.  The module is torch-only, so the arithmetic is
checked here on CPU tensors; tests/test_gpu_step.py runs the same function on the GPU through calc_metrics."""
import numpy as np
import pytest
import torch

import avsep_b200 as av
from oracle import bss_eval as BSS


def _case(B, N, L, seed):
    g = np.random.default_rng(seed)
    t = np.arange(L) / 11025.0
    ref = np.stack([[np.sin(2 * np.pi * 110.0 * (1 + n + 0.37 * b) * t) * (0.4 + 0.1 * n) + 0.02 * g.standard_normal(L)
                     for n in range(N)] for b in range(B)])
    est = ref + 0.2 * g.standard_normal(ref.shape) + 0.15 * ref[:, ::-1]          # noise + interference from the other sources
    return ref, est


@pytest.mark.parametrize("B,N,L", [(3, 2, 6000), (2, 3, 3000), (1, 1, 2048)])
def test_batch_matches_oracle(B, N, L):
    ref, est = _case(B, N, L, seed=B * 10 + N)
    out = av.bss_eval_batch(torch.from_numpy(ref), torch.from_numpy(est))
    for b in range(B):
        sdr, sir, sar = BSS.bss_eval_sources(ref[b], est[b])
        sis = BSS.si_sdr(ref[b], est[b])
        for key, want in (("sdr", sdr), ("sir", sir), ("sar", sar), ("si_sdr", sis)):
            if N == 1 and key == "sir":
                continue                                                          # no interferer: both are +inf / huge
            assert np.max(np.abs(out[key][b].numpy() - want)) < 1e-6, (key, b)    # dB; the parity bar is 0.01 dB


def test_float32_inputs_and_get_metrics_signature():
    ref, est = _case(1, 2, 5000, seed=5)
    md = av.bss_metrics.get_metrics(ref[0].sum(0), ref[0].astype(np.float32), est[0].astype(np.float32), sample_rate=11025,
                                    metrics_list=["sdr", "sir", "sar", "si_sdr"], device=torch.device("cpu"))
    want = BSS.get_metrics(None, ref[0].astype(np.float32), est[0].astype(np.float32))
    for k in ("sdr", "sir", "sar", "si_sdr"):
        assert abs(md[k] - want[k]) < 1e-6


def test_silent_reference_falls_back_to_least_squares():
    """mir_eval catches the singular Gram matrix of an all-zero reference and uses lstsq; so does the batch version."""
    ref, est = _case(1, 2, 3000, seed=8)
    ref[0, 1] = 0.0
    out = av.bss_eval_batch(torch.from_numpy(ref), torch.from_numpy(est))
    sdr, _, _ = BSS.bss_eval_sources(ref[0], est[0])
    assert torch.isfinite(out["sdr"][0, 0])
    assert abs(float(out["sdr"][0, 0]) - sdr[0]) < 0.01


def test_shape_errors():
    with pytest.raises(ValueError):
        av.bss_eval_batch(torch.zeros(2, 3000), torch.zeros(2, 3000))
