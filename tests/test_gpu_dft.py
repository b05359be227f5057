"""GPU parity of the tcgen05 DFT GEMM path (K1 STFT, K4 iSTFT / separate) against the CPU oracle.

Tolerance (BASELINE.json north_star): spectrogram / waveform max-norm relative error <= 1e-4.
The split-precision (bf16 x 3) GEMM is expected to land around 4e-6.
All calls go through the C-ABI (ctypes) of libavsep_b200.so.
"""
import os

import numpy as np
import pytest
import torch

from conftest import relerr
from oracle import np_oracle as O

import avsep_b200 as av

pytestmark = pytest.mark.gpu
TOL = 1e-4
L = 65535


def _signals(S, seed=1234, length=L):
    w = O.synth_sources(1, S, L=length, seed=seed)[0]          # [S, L], each already / S
    return np.ascontiguousarray(w * S * 0.5)


def _sync():
    torch.cuda.synchronize()


def _basis_expected(inverse):
    n = np.arange(1022)
    win = O.hann_periodic(1022)
    out = np.zeros((1024, 1024))
    R = np.arange(1024)
    m, q = R >> 7, (R >> 5) & 3
    if not inverse:
        part, i = (R >> 4) & 1, R & 15
        f = 64 * m + 16 * q + i
        ang = 2 * np.pi * ((f[:, None] * n[None, :]) % 1022) / 1022.0
        out[:, :1022] = np.where(part[:, None] == 0, np.cos(ang), -np.sin(ang)) * win[None, :]
    else:
        jj = R & 31
        nn = 32 * m + jj + 256 * q
        ok = nn < 1022
        f = np.arange(512)
        c = np.where((f == 0) | (f == 511), 1.0, 2.0)
        nv = np.where(ok, nn, 0)
        ang = 2 * np.pi * ((nv[:, None] * f[None, :]) % 1022) / 1022.0
        w = np.where(ok, win[nv], 0.0)[:, None]
        out[:, :512] = w * c[None, :] * np.cos(ang) / 1022.0
        out[:, 512:] = -w * c[None, :] * np.sin(ang) / 1022.0
    return out


def test_basis_tables():
    plan = av.Plan(1, L)
    for inverse in (0, 1):
        got = np.zeros((1024, 1024), dtype=np.float32)
        av._lib.check(av._lib.lib.avsep_debug_basis(plan.handle, inverse, got.ctypes.data))
        assert relerr(got, _basis_expected(inverse)) < 2e-5          # bf16 hi + lo = 16 significant bits
    plan.close()


def test_raw_gemm_matches_staged_operands():
    """Isolates the tensor-core pipeline of the plain CTA-pair GEMM (the in-library cross-check of the folded kernels):
    raw accumulators vs (hi+lo basis) @ (hi+lo staged frames) in float64."""
    plan = av.Plan(3, L)
    plan.set_fold(False)                  # the raw-accumulator hook exercises the plain 1024 x 1024 GEMM
    S = 3
    w = torch.from_numpy(_signals(S)).cuda()
    mag, _ = av.stft(w, plan=plan)
    _sync()
    T = plan.frames
    raw = torch.empty((S, 1024, T), device="cuda")
    av._lib.check(av._lib.lib.avsep_debug_raw_gemm(plan.handle, 0, S, raw.data_ptr(), torch.cuda.current_stream().cuda_stream))
    _sync()
    basis = np.zeros((1024, 1024), dtype=np.float32)
    av._lib.check(av._lib.lib.avsep_debug_basis(plan.handle, 0, basis.ctypes.data))
    for s in (0, S - 1):
        frames = np.zeros((T, 1024), dtype=np.float32)
        av._lib.check(av._lib.lib.avsep_debug_staged_frames(plan.handle, s, frames.ctypes.data))
        # staged operand = reflect-padded wave, 16 significant bits
        pad = np.pad(w[s].cpu().numpy().astype(np.float64), 511, mode="reflect")
        exp_frames = np.stack([pad[256 * t:256 * t + 1022] for t in range(T)])
        assert relerr(frames[:, :1022], exp_frames) < 2e-5
        expect = basis.astype(np.float64) @ frames.astype(np.float64).T
        assert relerr(raw[s].cpu().numpy(), expect) < 1e-5, f"signal {s}"
    plan.close()


def test_stft_vs_oracle():
    S = 3
    wav = _signals(S)
    mag, ph, re, im = av.stft(torch.from_numpy(wav).cuda(), complex_parts=True)
    _sync()
    assert mag.shape == (S, 512, 256) and ph.shape == mag.shape
    for s in range(S):
        spec = O.stft_complex(wav[s])
        amp_o, _ = O.stft_magphase(wav[s])
        assert relerr(mag[s].cpu().numpy(), amp_o) < TOL
        got = re[s].cpu().numpy().astype(np.float64) + 1j * im[s].cpu().numpy().astype(np.float64)
        assert np.max(np.abs(got - spec)) / np.max(np.abs(spec)) < TOL
        # phase is compared through the reconstructed spectrum (atan2 branch cut)
        rec = mag[s].cpu().numpy().astype(np.float64) * np.exp(1j * ph[s].cpu().numpy().astype(np.float64))
        assert np.max(np.abs(rec - spec)) / np.max(np.abs(spec)) < TOL
    err = relerr(mag[0].cpu().numpy(), O.stft_magphase(wav[0])[0])
    print(f"stft max-norm relative error {err:.3e}")


@pytest.mark.parametrize("length", [L, 512, 3000, 70001])
def test_folded_and_plain_gemm_agree(length):
    """The folded forward transform (operand generated in the kernel, dft_fold3) vs the plain DFT GEMM on the staged
    waveform vs the oracle -- including the reflect-padded edge frames, clips shorter than one tile and clips spanning
    two tiles -- and bit-reproducibility of the folded kernel."""
    S = 5
    wav = _signals(S, seed=21, length=length)
    w = torch.from_numpy(wav).cuda()
    plan = av.Plan(S, length)
    plan.set_fold(True)
    m1, p1, r1, i1 = av.stft(w, complex_parts=True, plan=plan)
    m2, p2, r2, i2 = av.stft(w, complex_parts=True, plan=plan)
    plan.set_fold(False)
    m0, p0, r0, i0 = av.stft(w, complex_parts=True, plan=plan)
    _sync()
    assert torch.equal(m1, m2) and torch.equal(r1, r2) and torch.equal(i1, i2) and torch.equal(p1, p2)
    scale = float(m0.max())
    assert float((m1 - m0).abs().max()) / scale < 2e-5
    assert float((r1 - r0).abs().max()) / scale < 2e-5 and float((i1 - i0).abs().max()) / scale < 2e-5
    for s in range(S):
        spec = O.stft_complex(wav[s])
        got = r1[s].cpu().numpy().astype(np.float64) + 1j * i1[s].cpu().numpy().astype(np.float64)
        e = np.max(np.abs(got - spec)) / np.max(np.abs(spec))
        assert e < TOL, e
    print(f"folded stft max-norm relative error {e:.3e}")
    plan.close()


@pytest.mark.parametrize("N,B,length", [(2, 3, L), (2, 2, 70001), (2, 1, 3000), (3, 2, L), (3, 3, 40000), (4, 2, L), (4, 1, 9000), (1, 2, L)])
def test_stft_mix_from_sources(N, B, length):
    """avsep_stft_mix: magnitudes of the mixtures and sources from the sources alone.  N = 2..4 take the linear route
    (dft_fold3m: |sum Xn| from the N source transforms, ragged clip lengths included); N = 1 and the phase output sum the
    waveform first."""
    g = np.random.default_rng(N * 100 + B)
    src = np.stack([_signals(B, seed=int(g.integers(1 << 20)), length=length) for _ in range(N)]) / N      # [N, B, L]
    mag, _ = av.stft_mix(torch.from_numpy(src).cuda())
    mixw = src.sum(axis=0).astype(np.float32)
    ref_mix, _ = av.stft(torch.from_numpy(mixw).cuda(), phase=False)
    ref_src, _ = av.stft(torch.from_numpy(src.reshape(N * B, length)).cuda(), phase=False)
    _sync()
    scale = float(ref_mix.max())
    assert float((mag[0] - ref_mix).abs().max()) / scale < 2e-5
    assert float((mag[1:].reshape(N * B, 512, -1) - ref_src).abs().max()) / scale < 2e-6
    for b in range(B):
        assert relerr(mag[0, b].cpu().numpy(), O.stft_magphase(mixw[b])[0]) < TOL
    if N in (2, 3):                                               # with the linear route switched off: the summed waveform's own transform
        plan = av.Plan((N + 1) * B, length)
        plan.set_mix_linear(False)
        m2, p2 = av.stft_mix(torch.from_numpy(src).cuda(), plan=plan)
        if N == 2:                                                # one IEEE add either way: same bits
            assert torch.equal(m2[0], ref_mix)
        else:
            assert float((m2[0] - ref_mix).abs().max()) / scale < 2e-6
        m3, p3 = av.stft_mix(torch.from_numpy(src).cuda(), phase=True, plan=plan)
        _sync()
        rec = m3[0, 0].cpu().numpy().astype(np.float64) * np.exp(1j * p3[0, 0].cpu().numpy().astype(np.float64))
        spec = O.stft_complex(mixw[0])
        assert np.max(np.abs(rec - spec)) / np.max(np.abs(spec)) < TOL
        plan.close()


def test_stft_golden_fixture(golden):
    g = golden("stft")
    src = O.synth_sources(1, 2, seed=1234)[0]
    mix = (src[0] + src[1]).astype(np.float32)
    mag, ph = av.stft(torch.from_numpy(np.stack([mix, src[0]])).cuda())
    fr = g["frames"]
    assert relerr(mag[0].cpu().numpy()[:, fr], g["amp_mix"]) < TOL
    assert relerr(mag[1].cpu().numpy()[:, fr], g["amp_src0"]) < TOL
    rec = mag[0].cpu().numpy()[:, fr].astype(np.float64) * np.exp(1j * ph[0].cpu().numpy()[:, fr].astype(np.float64))
    ref = g["amp_mix"].astype(np.float64) * np.exp(1j * g["phase_mix"].astype(np.float64))
    assert np.max(np.abs(rec - ref)) / np.max(np.abs(ref)) < TOL


@pytest.mark.parametrize("length", [3000, 512, 70001])
def test_stft_ragged_lengths(length, golden):
    """Edge cases: short clips (T < one tile), clips longer than one 256-frame tile."""
    wav = _signals(2, seed=5, length=length)
    mag, ph = av.stft(torch.from_numpy(wav).cuda())
    T = O.n_frames(length)
    assert mag.shape == (2, 512, T)
    for s in range(2):
        amp_o, _ = O.stft_magphase(wav[s])
        assert relerr(mag[s].cpu().numpy(), amp_o) < TOL
    if length == 3000:
        g = golden("stft")
        src = O.synth_sources(1, 2, seed=1234)[0]
        m2, _ = av.stft(torch.from_numpy(src[1][:3000].copy()).cuda())
        assert relerr(m2.cpu().numpy(), g["amp_short"]) < TOL


def test_stft_single_and_empty_batch():
    wav = _signals(1)
    m1, p1 = av.stft(torch.from_numpy(wav[0]).cuda())
    assert m1.shape == (512, 256)
    assert relerr(m1.cpu().numpy(), O.stft_magphase(wav[0])[0]) < TOL
    with pytest.raises(av.AvsepError):
        av.stft(torch.zeros(1, 300, device="cuda"))                 # shorter than the reflect padding
    with pytest.raises(av.AvsepError):
        av.stft(torch.zeros(2, L, device="cuda", dtype=torch.float64))


def test_istft_vs_oracle():
    wav = _signals(2, seed=9)
    amp, ph = zip(*[O.stft_magphase(w) for w in wav])
    amp, ph = np.stack(amp), np.stack(ph)
    mask = (np.random.default_rng(51).random((2, 512, 256)) > 0.4).astype(np.float32)
    out = av.istft(torch.from_numpy(amp * mask).cuda(), torch.from_numpy(ph).cuda())
    _sync()
    assert out.shape == (2, 65280)
    for s in range(2):
        ref = O.istft_reconstruction(amp[s] * mask[s], ph[s])
        e = relerr(out[s].cpu().numpy(), ref)
        assert e < TOL, e
    # clip to [-1, 1] (utils.py:104)
    loud = av.istft(torch.from_numpy(amp * 40).cuda(), torch.from_numpy(ph).cuda())
    assert float(loud.abs().max()) <= 1.0
    ref = O.istft_reconstruction(amp[0] * 40, ph[0])
    assert relerr(loud[0].cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("length", [65535, 33000, 5000, 65535 - 256 * 3])
def test_folded_inverse_equals_plain_inverse(length):
    """The folded inverse GEMM (idft_fold.cuh: two 512 x 512 contractions + mirrored overlap-add) against the plain
    1024 x 1024 inverse GEMM of the same library and against the fp64 oracle, at full and ragged frame counts."""
    S = 3
    wav = _signals(S, seed=21, length=length)
    amp, ph = zip(*[O.stft_magphase(w) for w in wav])
    amp, ph = np.stack(amp), np.stack(ph)
    T = amp.shape[-1]
    mask = (np.random.default_rng(5).random((S, 512, T)) > 0.3).astype(np.float32)
    m, p_ = torch.from_numpy(amp * mask).cuda(), torch.from_numpy(ph).cuda()
    pl = av.ops.Plan(S, length)
    try:
        assert pl.frames == T
        pl.set_inverse_fold(True)
        folded = av.istft(m, p_, plan=pl)
        folded2 = av.istft(m, p_, plan=pl)
        pl.set_inverse_fold(False)
        plain = av.istft(m, p_, plan=pl)
        _sync()
    finally:
        pl.close()
    assert folded.shape == (S, 256 * (T - 1))
    assert torch.equal(folded, folded2)                              # fixed summation order: bit-reproducible
    scale = float(plain.abs().max())
    assert float((folded - plain).abs().max()) / scale < 2e-5
    for s in range(S):
        ref = O.istft_reconstruction(amp[s] * mask[s], ph[s])
        assert relerr(folded[s].cpu().numpy(), ref) < TOL


def test_istft_golden_fixture(golden):
    g = golden("stft")
    src = O.synth_sources(1, 2, seed=1234)[0]
    mix = (src[0] + src[1]).astype(np.float32)
    amp, ph = O.stft_magphase(mix)
    mask = (np.random.default_rng(51).random((512, 256)) > 0.4).astype(np.float32)
    out = av.istft(torch.from_numpy((amp * mask)[None]).cuda(), torch.from_numpy(ph[None]).cuda())[0].cpu().numpy()
    assert out.shape[0] == int(g["istft_len"])
    scale = max(np.max(np.abs(g["istft_masked_head"])), np.max(np.abs(g["istft_masked_mid"])), np.max(np.abs(g["istft_masked_tail"])))
    assert np.max(np.abs(out[:4096] - g["istft_masked_head"])) / scale < TOL
    assert np.max(np.abs(out[30000:34096] - g["istft_masked_mid"])) / scale < TOL
    assert np.max(np.abs(out[-4096:] - g["istft_masked_tail"])) / scale < TOL


def test_istft_reconstruction_dropin():
    """utils.istft_reconstruction signature: numpy in, float64 numpy out."""
    wav = _signals(1, seed=3)[0]
    amp, ph = O.stft_magphase(wav)
    out = av.istft_reconstruction(amp, ph, hop_length=256)
    assert out.dtype == np.float64 and out.shape == (65280,)
    assert relerr(out, O.istft_reconstruction(amp, ph)) < TOL
    a2, p2 = av._stft(wav)
    assert a2.shape == (512, 256) and relerr(a2.cpu().numpy(), amp) < TOL


def test_mix_n_and_stft_reference_signature():
    """BaseDataset._mix_n_and_stft (dataset/base.py:174-189): list of N numpy sources + numpy mixture in,
    (amp_mix [1,512,T], mags list of [1,512,T], phase_mix [1,512,T]) out; `audios` entries become torch tensors."""
    N = 2
    src = O.synth_sources(1, N, seed=31)[0]                      # [N, L]
    mix = src.sum(axis=0).astype(np.float32)
    audios = [src[n].copy() for n in range(N)]
    amp_mix, mags, phase_mix = av._mix_n_and_stft(audios, mix)
    _sync()
    assert amp_mix.shape == (1, 512, 256) and phase_mix.shape == (1, 512, 256) and len(mags) == N
    assert all(torch.is_tensor(a) and not a.is_cuda and a.shape == (L,) for a in audios)      # base.py:186-187
    amp_o, ph_o = O.stft_magphase(mix)
    assert relerr(amp_mix[0].cpu().numpy(), amp_o) < TOL
    # the phase is ill-conditioned where the magnitude vanishes: compare the complex spectrum it reconstructs
    spec = amp_mix[0].cpu().numpy() * np.exp(1j * phase_mix[0].cpu().numpy().astype(np.float64))
    spec_o = amp_o * np.exp(1j * ph_o.astype(np.float64))
    assert np.max(np.abs(spec - spec_o)) / np.max(np.abs(spec_o)) < TOL
    for n in range(N):
        assert mags[n].shape == (1, 512, 256)
        assert relerr(mags[n][0].cpu().numpy(), O.stft_magphase(src[n])[0]) < TOL
    # phase values themselves where the bin carries energy
    strong = amp_o > 1e-2 * amp_o.max()
    d = np.angle(np.exp(1j * (phase_mix[0].cpu().numpy()[strong].astype(np.float64) - ph_o[strong])))
    assert np.max(np.abs(d)) < 1e-3


def test_stft_batch_collated_form():
    """stft_batch: collated [N,B,L] sources (+ optional mixture) -> the batch dict entries
    mag_mix [B,1,512,T], mags list, phase_mix, equal to the per-item _mix_n_and_stft results."""
    N, B = 2, 3
    src = O.synth_sources(B, N, seed=32)                         # [B, N, L]
    audios = torch.from_numpy(np.ascontiguousarray(src.transpose(1, 0, 2))).cuda()       # [N, B, L]
    out = av.stft_batch(audios)
    explicit = av.stft_batch(audios, audios.sum(dim=0))
    _sync()
    assert out["mag_mix"].shape == (B, 1, 512, 256) and out["phase_mix"].shape == (B, 1, 512, 256) and len(out["mags"]) == N
    for k in ("mag_mix", "phase_mix"):
        assert torch.equal(out[k], explicit[k])
    for b in range(B):
        mix = src[b].sum(axis=0).astype(np.float32)
        amp_o, ph_o = O.stft_magphase(mix)
        assert relerr(out["mag_mix"][b, 0].cpu().numpy(), amp_o) < TOL
        spec = out["mag_mix"][b, 0].cpu().numpy() * np.exp(1j * out["phase_mix"][b, 0].cpu().numpy().astype(np.float64))
        assert np.max(np.abs(spec - amp_o * np.exp(1j * ph_o.astype(np.float64)))) / np.max(amp_o) < TOL
        for n in range(N):
            assert relerr(out["mags"][n][b, 0].cpu().numpy(), O.stft_magphase(src[b, n])[0]) < TOL


def _block_masks(rng, N, B, H=256, T=256, lo=0.2, hi=0.9, blk=8):
    small = rng.random((N, B, H // blk, T // blk)) > 0.5
    m = np.where(small, hi, lo).astype(np.float32)
    return np.repeat(np.repeat(m, blk, axis=2), blk, axis=3)


@pytest.mark.parametrize("binary,log_freq", [(True, True), (False, True), (True, False)])
def test_separate_vs_oracle(binary, log_freq):
    B, N = 2, 2
    rng = np.random.default_rng(77)
    src = O.synth_sources(B, N, seed=77)
    mix = src.sum(axis=1).astype(np.float32)
    amp, ph = zip(*[O.stft_magphase(m) for m in mix])
    amp, ph = np.stack(amp)[:, None], np.stack(ph)[:, None]
    H = 256 if log_freq else 512
    if binary:
        masks = _block_masks(rng, N, B, H=H)
    else:
        masks = rng.random((N, B, H, 256)).astype(np.float32)
    pred = [torch.from_numpy(masks[n][:, None]).cuda() for n in range(N)]
    out = av.separate(pred, torch.from_numpy(amp).cuda(), torch.from_numpy(ph).cuda(), log_freq=log_freq,
                      binary_mask=binary, mask_thres=0.5)
    _sync()
    assert out.shape == (B, N + 1, 65280)
    mix_o, pred_o, _ = O.separate(amp, ph, [masks[n][:, None] for n in range(N)], log_freq=log_freq, binary_mask=binary)
    scale = np.max(np.abs(mix_o))
    got = out.cpu().numpy()
    assert np.max(np.abs(got[:, 0] - mix_o)) / scale < TOL
    for n in range(N):
        assert np.max(np.abs(got[:, 1 + n] - pred_o[:, n])) / scale < TOL


@pytest.mark.parametrize("binary,log_freq,N,B,length", [(True, True, 2, 2, L), (False, True, 3, 3, L), (True, False, 2, 1, L),
                                                         (False, True, 1, 2, 256 * 99 + 255),     # T = 100: ragged 32-frame tiles
                                                         (True, True, 4, 1, 256 * 43 + 255)])     # T = 44
def test_staged_operand_kernel_equals_register_tile_kernel(binary, log_freq, N, B, length):
    """spec_prep_tma_kernel (cp.async staging + TMA tensor stores) and spec_prep_vec_kernel build the SAME inverse operand:
    the separated waveforms are bit-identical; a 3-source batch at N = 3 also stages one source per pass on the low tile."""
    g = torch.Generator(device="cuda").manual_seed(5 + N)
    mix = (torch.rand((B, length), device="cuda", generator=g) - 0.5) * 0.9
    _, _, re, im = av.stft(mix, complex_parts=True)
    T = re.shape[-1]
    H = 256 if log_freq else 512
    masks = torch.rand((N, B, H, T), device="cuda", generator=g)
    plan = av.ops.Plan(B * (N + 1), length)
    outs = []
    for staged in (True, False):
        plan.set_spec_staged(staged)
        outs.append(av.separate(masks, None, None, re_mix=re, im_mix=im, log_freq=log_freq, binary_mask=binary, plan=plan).clone())
    _sync()
    assert outs[0].shape == (B, N + 1, 256 * (T - 1))
    assert torch.equal(outs[0], outs[1])
    assert float(outs[0].abs().max()) > 1e-3


def test_sm_budget_does_not_change_results():
    """avsep_plan_set_sm_budget only caps the persistent grids (tile order): forward and inverse transforms are bit-identical
    on 48 SMs and on the whole device."""
    B, N = 3, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    src = (torch.rand((N, B, L), device="cuda", generator=g) - 0.5) * 0.8
    plan = av.ops.Plan((N + 1) * B, L)
    masks = torch.rand((N, B, 256, 256), device="cuda", generator=g)
    outs = []
    for budget in (0, 48):
        plan.set_sm_budget(budget)
        mag, _ = av.stft_mix(src, plan=plan)
        m2, _, re, im = av.stft(src.sum(dim=0), complex_parts=True, plan=plan)
        wav = av.separate(masks, None, None, re_mix=re, im_mix=im, plan=plan)
        outs.append((mag.clone(), m2.clone(), wav.clone()))
    _sync()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    with pytest.raises(av.AvsepError):
        plan.set_sm_budget(1)


def test_separate_without_sources():
    """Edge case of the staged operand kernel: N = 0 (mixture reconstruction only, (re, im) input) equals the (mag, phase) inverse."""
    g = torch.Generator(device="cuda").manual_seed(21)
    mix = (torch.rand((2, L), device="cuda", generator=g) - 0.5) * 0.9
    mag, ph, re, im = av.stft(mix, complex_parts=True)
    none = torch.empty((0, 2, 256, mag.shape[-1]), device="cuda")
    a = av.separate(none, None, None, re_mix=re, im_mix=im)
    b = av.istft(mag, ph)
    _sync()
    assert a.shape == (2, 1, 65280)
    assert float((a[:, 0] - b).abs().max()) / float(b.abs().max()) < 2e-5
    assert float((a[:, 0] - mix[:, :65280]).abs().max()) < 1e-4          # iSTFT(STFT(x)) = x


def test_separate_with_complex_parts_equals_magphase():
    B, N = 1, 2
    src = O.synth_sources(B, N, seed=8)
    mix = torch.from_numpy(src.sum(axis=1).astype(np.float32)).cuda()
    mag, ph, re, im = av.stft(mix, complex_parts=True)
    masks = torch.rand((N, B, 256, 256), device="cuda")
    a = av.separate(masks, mag, ph, binary_mask=False)
    b = av.separate(masks, None, None, re_mix=re, im_mix=im, binary_mask=False)
    assert float((a - b).abs().max()) / float(a.abs().max()) < 2e-5


def test_roundtrip_and_linearity_at_full_batch():
    """Size-independent properties at BASELINE config sizes (B=80, N=2 -> 240 signals):
    iSTFT(STFT(x)) == x[:65280] and STFT(a + b) == STFT(a) + STFT(b)."""
    B, N = 80, 2
    g = torch.Generator(device="cuda").manual_seed(1234)
    src = (torch.rand((N, B, L), device="cuda", generator=g) - 0.5) * 0.8
    mix = src.sum(dim=0)
    waves = torch.cat([mix[None], src], dim=0).reshape((N + 1) * B, L)
    mag, ph, re, im = av.stft(waves, complex_parts=True)
    rec = av.istft(mag, ph)
    _sync()
    err = float((rec - waves[:, :65280]).abs().max())
    assert err < 1e-4 * float(waves.abs().max()) * 10, err          # fp32 round trip: ~1e-6 absolute
    assert err < 2e-5
    re = re.view(N + 1, B, 512, 256)
    im = im.view(N + 1, B, 512, 256)
    lin = max(float((re[0] - re[1] - re[2]).abs().max()), float((im[0] - im[1] - im[2]).abs().max()))
    assert lin / float(mag.max()) < 2e-5
    # determinism: same input, same bits
    mag2, _ = av.stft(waves, phase=False)
    assert torch.equal(mag, mag2)
    rec2 = av.istft(mag, ph)
    assert torch.equal(rec, rec2)
