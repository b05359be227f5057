"""The reference's own sequence of torch calls for the hot path, on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py header).  Used (i) to
cross-check the numpy oracle, and (ii) as the timed CPU arm of ``bench.py``
(``cpu_baseline`` and ``--impl reference``); ``tools/torch_gpu_baseline.py`` runs the same
calls with the tensors on ``cuda`` to report what stock PyTorch does on the B200.  /root/reference cannot travel to
the GPU box, so the handful of reference functions on the path are restated
here as the *same library calls the reference makes* (``F.grid_sample``,
``torch.bmm``, ``torch.sigmoid``, ``F.binary_cross_entropy``); ``torch.stft`` /
``torch.istft`` stand in for the un-vendored ``librosa.stft`` / ``istft``
(identical definition: periodic Hann, centre reflect padding, n_fft 1022).
``oracle/make_golden.py`` verifies each of them against the real reference
modules imported from /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

N_FFT, HOP = 1022, 256


def warpgrid(bs, HO, WO, warp=True):
    """utils.warpgrid (utils.py:12-26) -- rebuilt on the host every call, like the reference."""
    x = np.linspace(-1, 1, WO)
    y = np.linspace(-1, 1, HO)
    xv, yv = np.meshgrid(x, y)
    grid = np.zeros((bs, HO, WO, 2))
    if warp:
        grid_y = (np.power(21, (yv + 1) / 2) - 11) / 10
    else:
        grid_y = np.log(yv * 10 + 11) / np.log(21) * 2 - 1
    grid[:, :, :, 0] = xv
    grid[:, :, :, 1] = grid_y
    return grid.astype(np.float32)


def stft_magphase(wave: torch.Tensor):
    """dataset/base.py:142-147 with torch.stft standing in for librosa.stft.  wave [S, L] -> amp, phase [S,512,T]."""
    win = torch.hann_window(N_FFT, periodic=True, dtype=wave.dtype, device=wave.device)
    spec = torch.stft(wave, N_FFT, HOP, window=win, center=True, pad_mode="reflect", return_complex=True)
    return spec.abs(), spec.angle()


def istft_reconstruction(mag: torch.Tensor, phase: torch.Tensor):
    """utils.py:101-104 with torch.istft standing in for librosa.istft.  [S,512,T] -> [S, 256*(T-1)] float64."""
    spec = mag.to(torch.complex128) * torch.exp(1j * phase.to(torch.float64))
    win = torch.hann_window(N_FFT, periodic=True, dtype=torch.float64, device=mag.device)
    wav = torch.istft(spec, N_FFT, HOP, window=win, center=True)
    return wav.clamp(-1.0, 1.0)


def prepare(mag_mix, mags, log_freq=True, weighted_loss=True, binary_mask=True):
    """NetWrapper.prepare (main.py:51-95)."""
    mag_mix = mag_mix + 1e-10
    B, T = mag_mix.size(0), mag_mix.size(3)
    mags = list(mags)
    if log_freq:
        grid_warp = torch.from_numpy(warpgrid(B, 256, T, warp=True)).to(mag_mix.device)   # main.py:62-63
        mag_mix = F.grid_sample(mag_mix, grid_warp, align_corners=False)
        for n in range(len(mags)):
            mags[n] = F.grid_sample(mags[n], grid_warp, align_corners=False)
    if weighted_loss:
        weights = torch.clamp(torch.log1p(mag_mix), 1e-3, 10)
    else:
        weights = torch.ones_like(mag_mix)
    gt_masks = []
    for n in range(len(mags)):
        if binary_mask:
            gt_masks.append((mags[n] > 0.5 * mag_mix).float())
        else:
            g = mags[n] / mag_mix
            g.clamp_(0., 5.)
            gt_masks.append(g)
    log_mag_mix = torch.log(mag_mix).detach()
    return mags, mag_mix, log_mag_mix, gt_masks, weights


def innerprod(feat_img, feat_sound, scale, bias):
    """InnerProd.forward (models/synthesizer_net.py:12-19)."""
    sound_size = feat_sound.size()
    B, C = sound_size[0], sound_size[1]
    feat_img = feat_img.view(B, 1, C)
    z = torch.bmm(feat_img * scale, feat_sound.view(B, C, -1)).view(B, 1, *sound_size[2:])
    return z + bias


def bce_list(preds, targets, weight):
    """BCELoss list form (models/criterion.py:10-25, 44-49)."""
    errs = [F.binary_cross_entropy(p, t, weight=weight) for p, t in zip(preds, targets)]
    return torch.mean(torch.stack(errs))


def train_step(waves, feat_sound, vecs, scale, bias):
    """Config C2: STFT (N+1 signals) -> prepare -> synth + sigmoid -> weighted BCE -> backward.

    waves [B, N+1, L] (index 0 = mixture), feat_sound [B,K,256,256] requires_grad, vecs [B,N,K] requires_grad.
    Returns the scalar loss (float).
    """
    B, S, L = waves.shape
    amp, _ = stft_magphase(waves.reshape(B * S, L))
    amp = amp.view(B, S, 1, amp.shape[-2], amp.shape[-1])
    mag_mix = amp[:, 0]
    mags = [amp[:, 1 + n] for n in range(S - 1)]
    _, _, _, gt_masks, weights = prepare(mag_mix, mags)
    preds = [torch.sigmoid(innerprod(vecs[:, n, :], feat_sound, scale, bias)) for n in range(S - 1)]
    err = bce_list(preds, gt_masks, weights)
    err.backward()
    return float(err.detach())


def eval_step(waves, feat_sound, vecs, scale, bias, binary_mask=True, mask_thres=0.5):
    """Config C1/C3: STFT -> prepare -> synth + sigmoid -> un-warp -> mask -> iSTFT (no SDR)."""
    B, S, L = waves.shape
    with torch.no_grad():
        amp, phase = stft_magphase(waves.reshape(B * S, L))
        amp = amp.view(B, S, 1, amp.shape[-2], amp.shape[-1])
        phase = phase.view(B, S, 1, phase.shape[-2], phase.shape[-1])
        mag_mix, phase_mix = amp[:, 0], phase[:, 0]
        mags = [amp[:, 1 + n] for n in range(S - 1)]
        prepare(mag_mix, mags)
        outs = [istft_reconstruction(mag_mix[:, 0], phase_mix[:, 0])]
        for n in range(S - 1):
            p = torch.sigmoid(innerprod(vecs[:, n, :], feat_sound, scale, bias))
            grid_unwarp = torch.from_numpy(warpgrid(B, N_FFT // 2 + 1, p.size(3), warp=False)).to(p.device)   # main.py:216-219
            lin = F.grid_sample(p, grid_unwarp, align_corners=False)
            if binary_mask:
                lin = (lin > mask_thres).float()
            outs.append(istft_reconstruction(mag_mix[:, 0] * lin[:, 0], phase_mix[:, 0]))
    return torch.stack(outs, dim=1)
