"""CPU oracle for the mix-and-separate spectrogram hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU arm.  The product path (``avsep_b200``) never imports this module and
raises when its CUDA library is missing.

Every function is a plain numpy restatement (float64 arithmetic unless the
reference computes in float32) of one reference function; the docstring cites
the reference ``file:line`` it follows.  Third-party pieces the reference calls
but does not vendor are restated from their published algorithm:

* ``librosa.stft`` / ``librosa.istft`` (librosa < 0.10, unpinned by the
  reference; evidence: ``np.complex`` at utils.py:102, ``librosa.output`` at
  utils.py:219): periodic Hann of length n_fft, centre reflect padding,
  ``rfft`` per frame; inverse = ``irfft`` * window, overlap-add, divide by the
  window sum-square envelope where it is > tiny, trim n_fft//2 each side.
* ``asteroid.metrics.get_metrics`` -> ``pb_bss_eval`` -> ``mir_eval.separation.
  bss_eval_sources`` (BSS-eval v3, 512-tap filters) and ``pb_bss_eval``
  ``si_sdr``: see ``oracle/bss_eval.py``.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference's own code
imported from /root/reference (``utils.warpgrid``, ``InnerProd``, ``Bias``,
``BCELoss``/``L1Loss``/``L2Loss``, ``PitWrapper``) and against the torch
functions the reference calls (``F.grid_sample``, ``F.binary_cross_entropy``,
``torch.stft``/``torch.istft`` as the definitional stand-in for the absent
librosa).  ``oracle/make_golden.py`` generated ``tests/golden/*.npz`` that way;
``tests/test_oracle_golden.py`` re-checks the oracle against them on every run.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

N_FFT = 1022          # arguments.py: --stft_frame default
HOP = 256             # arguments.py: --stft_hop default
AUD_LEN = 65535       # arguments.py: --audLen default
N_BINS = N_FFT // 2 + 1


# ----------------------------------------------------------------------------
# STFT / iSTFT  (dataset/base.py:142-147, utils.py:101-104, librosa conventions)
# ----------------------------------------------------------------------------
def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) -- librosa's default window."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def n_frames(length: int, n_fft: int = N_FFT, hop: int = HOP) -> int:
    """Frame count of librosa.stft(center=True): 1 + (len + 2*(n_fft//2) - n_fft)//hop."""
    return 1 + (length + 2 * (n_fft // 2) - n_fft) // hop


def stft_complex(wave: np.ndarray, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """librosa.stft(audio, n_fft, hop_length) as called at dataset/base.py:143.

    wave: float[L] -> complex128[n_fft//2+1, T].  (librosa stores complex64; the
    caller casts.)  center=True, pad_mode='reflect', window = periodic Hann.
    """
    wave = np.asarray(wave, dtype=np.float64)
    pad = n_fft // 2
    y = np.pad(wave, pad, mode="reflect")
    T = 1 + (y.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(T)[None, :]
    frames = y[idx] * hann_periodic(n_fft)[:, None]
    return np.fft.rfft(frames, axis=0)


def stft_magphase(wave: np.ndarray, n_fft: int = N_FFT, hop: int = HOP):
    """BaseDataset._stft (dataset/base.py:142-147): amp = |spec|, phase = angle(spec), float32."""
    spec = stft_complex(wave, n_fft, hop).astype(np.complex64)
    return np.abs(spec).astype(np.float32), np.angle(spec).astype(np.float32)


def window_sumsquare(T: int, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """librosa.filters.window_sumsquare for the periodic Hann: sum_t w^2[n - t*hop]."""
    n = n_fft + hop * (T - 1)
    env = np.zeros(n, dtype=np.float64)
    w2 = hann_periodic(n_fft) ** 2
    for t in range(T):
        env[t * hop:t * hop + n_fft] += w2
    return env


def istft(spec: np.ndarray, hop: int = HOP) -> np.ndarray:
    """librosa.istft(spec, hop_length=hop): n_fft = 2*(F-1), center=True, length=None.

    spec: complex[F, T] -> float64[hop*(T-1)].
    """
    spec = np.asarray(spec, dtype=np.complex128)
    F, T = spec.shape
    n_fft = 2 * (F - 1)
    w = hann_periodic(n_fft)
    ytmp = np.fft.irfft(spec, n=n_fft, axis=0) * w[:, None]          # [n_fft, T]
    y = np.zeros(n_fft + hop * (T - 1), dtype=np.float64)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += ytmp[:, t]
    env = window_sumsquare(T, n_fft, hop)
    nz = env > np.finfo(np.float64).tiny
    y[nz] /= env[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def istft_reconstruction(mag: np.ndarray, phase: np.ndarray, hop_length: int = HOP) -> np.ndarray:
    """utils.istft_reconstruction (utils.py:101-104): mag*exp(1j*phase) -> istft -> clip(-1, 1)."""
    spec = mag.astype(np.complex128) * np.exp(1j * phase.astype(np.float64))
    wav = istft(spec, hop_length)
    return np.clip(wav, -1.0, 1.0)


# ----------------------------------------------------------------------------
# warpgrid + grid_sample  (utils.py:12-26; main.py:61-66, 216-220)
# ----------------------------------------------------------------------------
def warpgrid(bs: int, HO: int, WO: int, warp: bool = True) -> np.ndarray:
    """utils.warpgrid (utils.py:12-26): float32[bs, HO, WO, 2], x in ch0, y in ch1."""
    x = np.linspace(-1, 1, WO)
    y = np.linspace(-1, 1, HO)
    xv, yv = np.meshgrid(x, y)
    grid = np.zeros((bs, HO, WO, 2))
    if warp:
        gy = (np.power(21, (yv + 1) / 2) - 11) / 10
    else:
        gy = np.log(yv * 10 + 11) / np.log(21) * 2 - 1
    grid[:, :, :, 0] = xv
    grid[:, :, :, 1] = gy
    return grid.astype(np.float32)


def grid_sample(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False).

    Restates ATen's grid_sampler_2d CPU kernel in float32: unnormalise
    i = ((g + 1) * size - 1) / 2, taps floor(i), floor(i)+1, weights from the
    opposite corner, out-of-range taps contribute zero.
    x: float32[B, C, Hi, Wi], grid: float32[B, Ho, Wo, 2] -> float32[B, C, Ho, Wo].
    """
    x = np.asarray(x, dtype=np.float32)
    grid = np.asarray(grid, dtype=np.float32)
    B, C, Hi, Wi = x.shape
    gx = grid[..., 0]
    gy = grid[..., 1]
    ix = ((gx + np.float32(1)) * np.float32(Wi) - np.float32(1)) / np.float32(2)
    iy = ((gy + np.float32(1)) * np.float32(Hi) - np.float32(1)) / np.float32(2)
    ix_nw = np.floor(ix)
    iy_nw = np.floor(iy)
    ix_se = ix_nw + 1
    iy_se = iy_nw + 1
    nw = (ix_se - ix) * (iy_se - iy)
    ne = (ix - ix_nw) * (iy_se - iy)
    sw = (ix_se - ix) * (iy - iy_nw)
    se = (ix - ix_nw) * (iy - iy_nw)
    out = np.zeros((B, C) + gx.shape[1:], dtype=np.float32)
    bidx = np.arange(B)[:, None, None]
    for (xx, yy, wgt) in ((ix_nw, iy_nw, nw), (ix_se, iy_nw, ne), (ix_nw, iy_se, sw), (ix_se, iy_se, se)):
        xi = xx.astype(np.int64)
        yi = yy.astype(np.int64)
        ok = (xi >= 0) & (xi < Wi) & (yi >= 0) & (yi < Hi)
        xc = np.clip(xi, 0, Wi - 1)
        yc = np.clip(yi, 0, Hi - 1)
        for c in range(C):
            v = x[bidx, c, yc, xc]
            out[:, c] += np.where(ok, v * wgt, np.float32(0)).astype(np.float32)
    return out


def axis_taps(n_out: int, n_in: int, coords: np.ndarray):
    """Separable form of the sampling above: per output index the lower tap and its fraction.

    coords: normalised grid coordinate per output index (float32, as warpgrid stores it).
    Returns (i0 int32[n_out], frac float32[n_out]); taps i0, i0+1 with weights (1-frac, frac),
    out-of-range taps are dropped by the consumer.
    """
    g = np.asarray(coords, dtype=np.float32)
    i = ((g + np.float32(1)) * np.float32(n_in) - np.float32(1)) / np.float32(2)
    i0 = np.floor(i)
    return i0.astype(np.int32), (i - i0).astype(np.float32)


def warp_axis_coords(HO: int, warp: bool) -> np.ndarray:
    """The y column of warpgrid (utils.py:19-22) as a 1-D float32 vector."""
    y = np.linspace(-1, 1, HO)
    if warp:
        gy = (np.power(21, (y + 1) / 2) - 11) / 10
    else:
        gy = np.log(y * 10 + 11) / np.log(21) * 2 - 1
    return gy.astype(np.float32)


def log_freq_warp(x: np.ndarray, HO: int, warp: bool) -> np.ndarray:
    """grid_sample(x, warpgrid(B, HO, W, warp)) -- the call at main.py:62-66 / 216-220."""
    B, _, _, W = x.shape
    return grid_sample(x, warpgrid(B, HO, W, warp))


# ----------------------------------------------------------------------------
# NetWrapper.prepare  (main.py:51-95)
# ----------------------------------------------------------------------------
def prepare(mag_mix: np.ndarray, mags, *, log_freq=True, weighted_loss=True, binary_mask=True):
    """NetWrapper.prepare (main.py:51-95), tensor part only.

    mag_mix float32[B,1,F,T]; mags: list of N float32[B,1,F,T].
    Returns dict(mags=list, mag_mix, log_mag_mix, gt_masks=list, weights), all float32[B,1,256,T].
    """
    f32 = np.float32
    mag_mix = (mag_mix.astype(f32) + f32(1e-10)).astype(f32)
    mags = [m.astype(f32) for m in mags]
    if log_freq:
        mag_mix = log_freq_warp(mag_mix, 256, True)
        mags = [log_freq_warp(m, 256, True) for m in mags]
    if weighted_loss:
        weights = np.clip(np.log1p(mag_mix.astype(np.float64)).astype(f32), f32(1e-3), f32(10))
    else:
        weights = np.ones_like(mag_mix)
    gt = []
    for m in mags:
        if binary_mask:
            gt.append((m > f32(0.5) * mag_mix).astype(f32))
        else:
            gt.append(np.clip((m / mag_mix).astype(f32), f32(0), f32(5)))
    log_mag_mix = np.log(mag_mix.astype(np.float64)).astype(f32)
    return dict(mags=mags, mag_mix=mag_mix, log_mag_mix=log_mag_mix, gt_masks=gt, weights=weights)


# ----------------------------------------------------------------------------
# Synthesizer  (models/synthesizer_net.py:6-70) and activation (models/__init__.py:16-28)
# ----------------------------------------------------------------------------
def innerprod_forward(feat_img, feat_sound, scale=None, bias=0.0):
    """InnerProd.forward (synthesizer_net.py:12-19); scale=None gives Bias.forward (:46-51).

    feat_img [B,K], feat_sound [B,K,H,W] -> z float64 [B,1,H,W] (cast by the caller).
    """
    fi = np.asarray(feat_img, dtype=np.float64)
    fs = np.asarray(feat_sound, dtype=np.float64)
    if scale is not None:
        fi = fi * np.asarray(scale, dtype=np.float64)[None, :]
    z = np.einsum("bk,bkhw->bhw", fi, fs)[:, None] + float(np.asarray(bias).reshape(-1)[0])
    return z


def innerprod_forward_nosum(feat_img, feat_sound, scale=None, bias=0.0):
    """InnerProd.forward_nosum (synthesizer_net.py:21-26): per-channel product [B,K,H,W]."""
    fi = np.asarray(feat_img, dtype=np.float64)
    if scale is not None:
        fi = fi * np.asarray(scale, dtype=np.float64)[None, :]
    return fi[:, :, None, None] * np.asarray(feat_sound, dtype=np.float64) + float(np.asarray(bias).reshape(-1)[0])


def innerprod_forward_pixelwise(feats_img, feat_sound, scale=None, bias=0.0):
    """InnerProd.forward_pixelwise (synthesizer_net.py:29-38): [B,K,HI,WI] x [B,K,HS,WS] -> [B,HI,WI,HS,WS]."""
    fi = np.asarray(feats_img, dtype=np.float64)
    if scale is not None:
        fi = fi * np.asarray(scale, dtype=np.float64)[None, :, None, None]
    return np.einsum("bkij,bkhw->bijhw", fi, np.asarray(feat_sound, dtype=np.float64)) \
        + float(np.asarray(bias).reshape(-1)[0])


def sigmoid(z):
    z = np.asarray(z, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-z))


def sigmoid_f32(z):
    """torch.sigmoid on float32 input: 1 / (1 + exp(-z)) evaluated IN float32, as ATen does.

    The float32 evaluation matters: for z >~ 16.7 the sum 1 + exp(-z) rounds to 1 and the result
    saturates to exactly 1.0 (a correctly rounded float64 sigmoid would give 0.99999994), which is
    what makes the reference's BCE term clamp to 100 * w there.
    """
    z = np.asarray(z, dtype=np.float32)
    with np.errstate(over="ignore"):
        e = np.exp(-z)
    return (np.float32(1) / (np.float32(1) + e)).astype(np.float32)


# ----------------------------------------------------------------------------
# Criterion  (models/criterion.py:6-49)
# ----------------------------------------------------------------------------
def _bce_elem(p, t):
    """F.binary_cross_entropy element: -(t*max(log p,-100) + (1-t)*max(log(1-p),-100))."""
    p = np.asarray(p, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p), -100.0)
        l1p = np.maximum(np.log1p(-p), -100.0)
    return -(t * lp + (1.0 - t) * l1p)


def loss_elem(kind: str, p, t):
    p64 = np.asarray(p, dtype=np.float64)
    t64 = np.asarray(t, dtype=np.float64)
    if kind == "bce":
        return _bce_elem(p64, t64)
    if kind == "l1":
        return np.abs(p64 - t64)
    if kind == "l2":
        return (p64 - t64) ** 2
    raise ValueError(kind)


def criterion(kind: str, preds, targets, weight=None) -> float:
    """BaseLoss.forward (criterion.py:10-25): list form = mean_n(mean(w * l(pred_n, tgt_n)))."""
    if isinstance(preds, (list, tuple)):
        w = 1.0 if weight is None else np.asarray(weight, dtype=np.float64)
        errs = [np.mean(w * loss_elem(kind, p, t)) for p, t in zip(preds, targets)]
        return float(np.mean(errs))
    w = 1.0 if weight is None else np.asarray(weight, dtype=np.float64)
    return float(np.mean(w * loss_elem(kind, preds, targets)))


def loss_grad_wrt_pred(kind: str, p, t, w, numel: int, nsrc: int):
    """d criterion / d pred_n as torch autograd produces it (upstream grad 1).

    BCE: w*(p-t)/max(p*(1-p), 1e-12) / numel / nsrc  (ATen binary_cross_entropy_backward, EPSILON=1e-12).
    """
    p = np.asarray(p, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    if kind == "bce":
        g = w * (p - t) / np.maximum(p * (1.0 - p), 1e-12)
    elif kind == "l1":
        g = w * np.sign(p - t)
    elif kind == "l2":
        g = w * 2.0 * (p - t)
    else:
        raise ValueError(kind)
    return g / float(numel) / float(nsrc)


def synth_loss_and_grads(kind, feat_sound, vecs, scale, bias, gts, weight, activation="sigmoid"):
    """Forward + analytic backward of the SoP++ synthesizer step (SoP++/main.py:203-209):

        for n: pred_n = sigmoid(InnerProd(vecs[:, n], feat_sound));  err = crit(preds, gts, weight)

    feat_sound [B,K,H,W] f32, vecs [B,N,K] f32, scale [K] or None (Bias), bias scalar,
    gts list of N [B,1,H,W], weight [B,1,H,W] or None.
    Returns dict(loss, preds(list, f32), dfeat [B,K,H,W], dvecs [B,N,K], dscale [K], dbias) in float64,
    following torch semantics: sigmoid output is rounded to float32 before the loss, so the
    gradient is exactly zero where it saturates to 0 or 1.
    """
    fs = np.asarray(feat_sound, dtype=np.float64)
    B, K, H, W = fs.shape
    vecs = np.asarray(vecs, dtype=np.float64)
    N = vecs.shape[1]
    sc = np.ones(K) if scale is None else np.asarray(scale, dtype=np.float64)
    b0 = float(np.asarray(bias).reshape(-1)[0])
    w = np.ones((B, 1, H, W)) if weight is None else np.asarray(weight, dtype=np.float64)
    numel = B * H * W
    preds, loss_terms = [], []
    dfeat = np.zeros_like(fs)
    dvecs = np.zeros((B, N, K))
    dscale = np.zeros(K)
    dbias = 0.0
    for n in range(N):
        z = np.einsum("bk,bkhw->bhw", vecs[:, n] * sc[None], fs)[:, None] + b0
        z32 = z.astype(np.float32)
        if activation == "sigmoid":
            p32 = sigmoid_f32(z32)
        else:
            p32 = z32
        preds.append(p32)
        p = p32.astype(np.float64)
        loss_terms.append(np.mean(w * loss_elem(kind, p, gts[n])))
        dp = loss_grad_wrt_pred(kind, p, gts[n], w, numel, N)
        dz = dp * p * (1.0 - p) if activation == "sigmoid" else dp
        dz = dz[:, 0]                                              # [B,H,W]
        dfeat += (vecs[:, n] * sc[None])[:, :, None, None] * dz[:, None]
        s = np.einsum("bhw,bkhw->bk", dz, fs)                      # sum_p dz * feat
        dvecs[:, n] = s * sc[None]
        dscale += np.sum(s * vecs[:, n], axis=0)
        dbias += float(np.sum(dz))
    return dict(loss=float(np.mean(loss_terms)), preds=preds, dfeat=dfeat, dvecs=dvecs,
                dscale=dscale, dbias=dbias)


# ----------------------------------------------------------------------------
# PitWrapper  (models/criterion.py:74-231)
# ----------------------------------------------------------------------------
def pit(kind: str, preds, targets, weights):
    """PitWrapper(F.binary_cross_entropy).forward (criterion.py:202-231).

    preds/targets/weights float[B, H, W, S].  loss_mat[i, j] = mean_{HW} w_i * l(pred_j, tgt_i)
    (criterion.py:159-176: pred repeated along a new axis -2, target/weight along -1), then the
    minimum over permutations p of mean_i loss_mat[i, p_i] in itertools order, first minimum wins
    (criterion.py:129-136).  Returns (loss float64[B], perms list[tuple]).
    """
    preds = np.asarray(preds, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    B = preds.shape[0]
    S = preds.shape[-1]
    losses, perms = [], []
    for b in range(B):
        p = preds[b][..., None, :]            # [..., 1, S]  (pred index j on the last axis)
        t = targets[b][..., :, None]          # [..., S, 1]  (target index i on axis -2)
        w = weights[b][..., :, None]
        lm = (w * loss_elem(kind, np.broadcast_to(p, p.shape[:-2] + (S, S)),
                            np.broadcast_to(t, t.shape[:-2] + (S, S))))
        lm = lm.reshape(-1, S, S).mean(axis=0)
        best, best_p = None, None
        for perm in itertools.permutations(range(S)):
            c = float(np.mean(lm[np.arange(S), list(perm)]))
            if best is None or best > c:
                best, best_p = c, perm
        losses.append(best)
        perms.append(best_p)
    return np.asarray(losses), perms


def pit_grad(kind: str, preds, targets, weights, perms, gloss):
    """Gradient of sum_b gloss[b] * pit_loss[b] w.r.t. preds (autograd of criterion.py:159-176 with the chosen
    permutation): prediction j is matched with target i where perms[b][i] == j."""
    preds = np.asarray(preds, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    B, S = preds.shape[0], preds.shape[-1]
    P = preds[0].size // S
    out = np.zeros_like(preds)
    for b in range(B):
        for i, j in enumerate(perms[b]):
            p, t, w = preds[b][..., j], targets[b][..., i], weights[b][..., i]
            if kind == "bce":
                g = w * (p - t) / np.maximum(p * (1 - p), 1e-12)
            elif kind == "l1":
                g = w * np.sign(p - t)
            else:
                g = w * 2 * (p - t)
            out[b][..., j] = gloss[b] * g / (S * P)
    return out


def pit_reorder(tensor, perms):
    """PitWrapper.reorder_tensor (criterion.py:180-200): out[b] = tensor[b][..., perms[b]]."""
    out = np.zeros_like(tensor)
    for b in range(tensor.shape[0]):
        out[b] = tensor[b][..., list(perms[b])]
    return out


# ----------------------------------------------------------------------------
# calc_metrics front half  (main.py:197-249): un-warp, threshold, mask, iSTFT
# ----------------------------------------------------------------------------
def separate(mag_mix, phase_mix, pred_masks, *, log_freq=True, binary_mask=True, mask_thres=0.5,
             n_bins=N_BINS, hop=HOP):
    """The reconstruction half of calc_metrics (main.py:213-249).

    mag_mix, phase_mix float32[B,1,F,T] (the un-warped, un-epsilon'd CPU copies, main.py:206-207);
    pred_masks list of N float32[B,1,256,T].
    Returns (mix_wav float64[B, L], preds_wav float64[B, N, L], masks_linear list of [B,1,F,T]).
    The all-zero guard (main.py:248-249, unseeded RNG) is NOT applied here; callers handle it.
    """
    B = mag_mix.shape[0]
    N = len(pred_masks)
    lin = []
    for n in range(N):
        m = np.asarray(pred_masks[n], dtype=np.float32)
        if log_freq:
            m = log_freq_warp(m, n_bins, False)
        if binary_mask:
            m = (m > np.float32(mask_thres)).astype(np.float32)
        lin.append(m)
    L = hop * (mag_mix.shape[3] - 1)
    mix = np.zeros((B, L))
    out = np.zeros((B, N, L))
    for j in range(B):
        mix[j] = istft_reconstruction(mag_mix[j, 0], phase_mix[j, 0], hop)
        for n in range(N):
            out[j, n] = istft_reconstruction(mag_mix[j, 0] * lin[n][j, 0], phase_mix[j, 0], hop)
    return mix, out, lin


# ----------------------------------------------------------------------------
# Synthetic MUSIC-shaped inputs; shared by tests and bench
# ----------------------------------------------------------------------------
def synth_sources(B: int, N: int, L: int = AUD_LEN, sr: int = 11025, seed: int = 1234) -> np.ndarray:
    """Harmonic-tone sources float32[B, N, L], already divided by N (dataset/music.py:120)."""
    rng = np.random.default_rng(seed)
    t = np.arange(L, dtype=np.float64) / sr
    out = np.zeros((B, N, L), dtype=np.float32)
    for b in range(B):
        for n in range(N):
            f0 = math.exp(rng.uniform(math.log(110.0), math.log(880.0)))
            sig = np.zeros(L)
            for h in range(1, 9):
                sig += (1.0 / h) * np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi))
            sig *= 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.2, 2.0) * t)
            sig += 0.01 * rng.standard_normal(L)
            sig *= rng.uniform(0.5, 1.0) / np.max(np.abs(sig))
            sig = np.clip(sig, -1.0, 1.0) / N
            out[b, n] = sig.astype(np.float32)
    return out
