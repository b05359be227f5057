"""The separation-metric evaluators against the CPU restatement of mir_eval / pb_bss_eval in
oracle/bss_eval.py: bss_eval_batch_torch (library-built cross-check, checked here on CPU tensors) and, on the GPU,
bss_eval_batch = the hand-written kernels of csrc/bss_eval.cu (direct correlations + block-Toeplitz Levinson)."""
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
    out = av.bss_eval_batch_torch(torch.from_numpy(ref), torch.from_numpy(est))
    for b in range(B):
        sdr, sir, sar = BSS.bss_eval_sources(ref[b], est[b])
        sis = BSS.si_sdr(ref[b], est[b])
        for key, want in (("sdr", sdr), ("sir", sir), ("sar", sar), ("si_sdr", sis)):
            if N == 1 and key == "sir":
                continue                                                          # no interferer: both are +inf / huge
            assert np.max(np.abs(out[key][b].numpy() - want)) < 1e-6, (key, b)    # dB; the parity bar is 0.01 dB


def test_float32_inputs_and_get_metrics_signature():
    ref, est = _case(1, 2, 5000, seed=5)
    out = av.bss_eval_batch_torch(torch.from_numpy(ref.astype(np.float32)), torch.from_numpy(est.astype(np.float32)))
    want = BSS.get_metrics(None, ref[0].astype(np.float32), est[0].astype(np.float32))
    for k in ("sdr", "sir", "sar", "si_sdr"):
        assert abs(float(out[k][0].mean()) - want[k]) < 1e-6


def test_silent_reference_falls_back_to_least_squares():
    """mir_eval catches the singular Gram matrix of an all-zero reference and uses lstsq; so does the batch version."""
    ref, est = _case(1, 2, 3000, seed=8)
    ref[0, 1] = 0.0
    out = av.bss_eval_batch_torch(torch.from_numpy(ref), torch.from_numpy(est))
    sdr, _, _ = BSS.bss_eval_sources(ref[0], est[0])
    assert torch.isfinite(out["sdr"][0, 0])
    assert abs(float(out["sdr"][0, 0]) - sdr[0]) < 0.01


def test_shape_errors():
    with pytest.raises(ValueError):
        av.bss_eval_batch_torch(torch.zeros(2, 3000), torch.zeros(2, 3000))
    with pytest.raises(ValueError):
        av.bss_eval_batch(torch.zeros(2, 3000), torch.zeros(2, 3000))
    with pytest.raises(av.AvsepError):                                            # the product evaluator has no CPU path
        av.bss_eval_batch(torch.zeros(1, 2, 3000), torch.zeros(1, 2, 3000))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,L,dtype", [(3, 2, 6000, torch.float64), (2, 3, 3000, torch.float32), (1, 1, 2048, torch.float32),
                                         (2, 2, 65280, torch.float32), (1, 3, 65280, torch.float32), (1, 4, 9000, torch.float64)])
def test_kernel_evaluator_matches_oracle(B, N, L, dtype):
    """avsep_bss_eval (direct fp64 correlations, block Levinson, quadratic-form energies) against the restatement of
    mir_eval / pb_bss_eval: within 1e-6 dB (the north-star bar is 0.01 dB), at small and at full clip length."""
    ref, est = _case(B, N, L, seed=B * 10 + N)
    tr, te = torch.from_numpy(ref).to(dtype), torch.from_numpy(est).to(dtype)
    out = av.bss_eval_batch(tr.cuda(), te.cuda())
    out2 = av.bss_eval_batch(tr.cuda(), te.cuda())
    for b in range(B):
        sdr, sir, sar = BSS.bss_eval_sources(tr[b].numpy(), te[b].numpy())
        sis = BSS.si_sdr(tr[b].numpy().astype(np.float64), te[b].numpy().astype(np.float64))
        for key, want in (("sdr", sdr), ("sir", sir), ("sar", sar), ("si_sdr", sis)):
            assert torch.equal(out[key], out2[key])                               # fixed summation orders: bit-reproducible
            if N == 1 and key == "sir":
                continue
            assert np.max(np.abs(out[key][b].cpu().numpy() - want)) < 1e-6, (key, b, out[key][b], want)


@pytest.mark.gpu
def test_kernel_evaluator_strided_views_and_library_cross_check():
    """The evaluator scores strided views in place (the step's waveform layouts) and agrees with the library-built
    torch.fft / torch.linalg route on a larger batch of MUSIC-shaped clips."""
    import bench
    B, N = 16, 2
    waves = bench.make_waves(B, N, 99).cuda()                       # [N+1, B, L] signal-major
    g = torch.Generator(device="cuda").manual_seed(3)
    wav = torch.empty((B, N + 1, 65280), device="cuda")
    mixm = torch.tensor([[1.0, 0.3], [0.2, 0.9]], device="cuda")
    wav[:, 1:] = torch.einsum("ij,jbl->bil", mixm, waves[1:, :, :65280]) + 0.02 * torch.randn((B, N, 65280), device="cuda", generator=g)
    refs_view = waves[1:, :, :65280].permute(1, 0, 2)               # [B, N, L] strided view, no copy
    ests_view = wav[:, 1:]
    got = av.bss_eval_batch(refs_view, ests_view)
    lib = av.bss_eval_batch_torch(refs_view.contiguous(), ests_view.contiguous())
    for k in ("sdr", "sir", "sar", "si_sdr"):
        assert float((got[k] - lib[k]).abs().max()) < 1e-6, k

