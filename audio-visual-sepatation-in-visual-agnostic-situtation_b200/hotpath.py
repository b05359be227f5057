"""Host-side mirror of the reference's functions on the hot path (same names, argument meaning and
error behaviour), so main.py / SoP++/main.py / inference.py / dataset/base.py can switch to the
B200 path by changing imports only.  See INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib, bss_metrics, ops
from ._lib import check, lib


# ------------------------------------------------------------------------------------------------
# utils.py
# ------------------------------------------------------------------------------------------------
def warpgrid(bs, HO, WO, warp=True):
    """utils.warpgrid (utils.py:12-26), kept for call-site compatibility only.

    The B200 path never materialises this grid: `warp()` below applies the same resample from two
    tap tables.  Code that still calls F.grid_sample(x, warpgrid(...)) gets the identical grid.
    """
    x = np.linspace(-1, 1, WO)
    y = np.linspace(-1, 1, HO)
    xv, yv = np.meshgrid(x, y)
    grid = np.zeros((bs, HO, WO, 2))
    grid_x = xv
    if warp:
        grid_y = (np.power(21, (yv + 1) / 2) - 11) / 10
    else:
        grid_y = np.log(yv * 10 + 11) / np.log(21) * 2 - 1
    grid[:, :, :, 0] = grid_x
    grid[:, :, :, 1] = grid_y
    return grid.astype(np.float32)


def warp(x, HO, warp=True):
    """Replaces `F.grid_sample(x, torch.from_numpy(warpgrid(B, HO, T, warp)).to(device), align_corners=False)`
    (main.py:62-66, 216-220; inference.py:47-49, 458-462; SoP++/main.py:57-61, 283-285)."""
    return ops.warp(x, HO, warp)


def istft_reconstruction(mag, phase, hop_length=256):
    """utils.istft_reconstruction (utils.py:101-104): numpy [512,T] x2 -> numpy float64 [256 (T-1)] in [-1, 1]."""
    if hop_length != ops.HOP:
        raise _lib.AvsepError(f"istft_reconstruction: only hop_length={ops.HOP} is supported, got {hop_length}")
    m = torch.as_tensor(np.ascontiguousarray(mag), dtype=torch.float32).cuda(non_blocking=True)
    p = torch.as_tensor(np.ascontiguousarray(phase), dtype=torch.float32).cuda(non_blocking=True)
    wav = ops.istft(m[None], p[None])[0]
    return wav.double().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# dataset/base.py
# ------------------------------------------------------------------------------------------------
def _stft(audio, stft_frame=1022, stft_hop=256, device="cuda"):
    """BaseDataset._stft (dataset/base.py:142-147): numpy float32 [L] -> (amp, phase) torch [512, T].

    Runs on the GPU; call it from the main process on collated waveforms, not from forked
    DataLoader workers (SURVEY.md section 8(b) threading)."""
    if stft_frame != ops.N_FFT or stft_hop != ops.HOP:
        raise _lib.AvsepError(f"_stft: only n_fft={ops.N_FFT}, hop={ops.HOP} are supported")
    w = torch.as_tensor(np.ascontiguousarray(audio), dtype=torch.float32).to(device, non_blocking=True)
    amp, phase = ops.stft(w)
    return amp, phase


def _mix_n_and_stft(audios, audio_mix, device="cuda"):
    """BaseDataset._mix_n_and_stft (dataset/base.py:174-189): one batched STFT for the mix and the N sources.

    audios: list of N numpy [L]; audio_mix numpy [L] -> (amp_mix [1,512,T], mags list of [1,512,T], phase_mix [1,512,T]);
    like the reference, `audios` entries are replaced by torch tensors."""
    N = len(audios)
    mix = torch.from_numpy(np.ascontiguousarray(audio_mix, dtype=np.float32)[None]).to(device, non_blocking=True)
    amp_mix, phase_mix = ops.stft(mix, phase=True)                 # |.| and angle from the kernel's epilogue (base.py:145-146)
    mags = []
    if N:
        src = torch.from_numpy(np.stack([np.asarray(a, dtype=np.float32) for a in audios])).to(device, non_blocking=True)
        amp_src, _ = ops.stft(src, phase=False)                    # the reference discards the sources' phase (base.py:181)
        mags = [amp_src[n].unsqueeze(0) for n in range(N)]
    for n in range(N):
        audios[n] = torch.from_numpy(np.asarray(audios[n]))
    return amp_mix[0].unsqueeze(0), mags, phase_mix[0].unsqueeze(0)


def stft_batch(audios: torch.Tensor, audio_mix: Optional[torch.Tensor] = None):
    """Batched form for the collated batch: audios [N,B,L] (+ audio_mix [B,L],
    default sum over sources, dataset/music.py:127) -> dict(mag_mix [B,1,512,T], mags list, phase_mix)."""
    N, B, L = audios.shape
    mix = audios.sum(dim=0) if audio_mix is None else audio_mix
    w = torch.cat([mix[None], audios], dim=0).reshape((N + 1) * B, L)
    mag, phase = ops.stft(w, phase=True)
    mag = mag.view(N + 1, B, 1, ops.N_BINS, -1)
    phase = phase.view(N + 1, B, 1, ops.N_BINS, -1)
    return dict(mag_mix=mag[0], mags=[mag[1 + n] for n in range(N)], phase_mix=phase[0])


# ------------------------------------------------------------------------------------------------
# main.py
# ------------------------------------------------------------------------------------------------
def prepare(batch_data, args, use_vis=False, is_share=False, load_clips=False):
    """NetWrapper.prepare (main.py:51-95) with the same return tuples; like the reference it rebinds
    batch_data['mags'][n] to the warped magnitudes (main.py:66)."""
    mag_mix = batch_data['mag_mix']
    mags = batch_data['mags']
    N = args.num_mix
    mags_w, mix_w, log_mix, gt_masks, weights = ops.prepare(
        mag_mix, [mags[n] for n in range(N)], log_freq=bool(args.log_freq),
        weighted_loss=bool(args.weighted_loss), binary_mask=bool(args.binary_mask))
    for n in range(N):
        mags[n] = mags_w[n]
    if use_vis or getattr(args, "fusion_type", None) == "share" or is_share:
        frames = batch_data['frames']
        clips = batch_data['clips'] if load_clips else None
        return frames, clips, mags, mix_w, log_mix, gt_masks, weights
    return mags, mix_w, log_mix, gt_masks, weights


class _Meter:
    """utils.AverageMeter (utils.py:39-78), unweighted."""

    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, val):
        self.sum += float(val)
        self.count += 1

    def average(self):
        return self.sum / self.count if self.count else 0.0


def calc_metrics(batch_data, outputs, args, get_metrics: Optional[Callable] = None, rng: Optional[np.random.Generator] = None):
    """calc_metrics (main.py:197-286): un-warp, threshold, mask, iSTFT on the GPU in one call for the
    whole batch, then the per-clip metric call.  `get_metrics` defaults to asteroid's (as the reference
    imports it, main.py:25); pass any callable with the same signature, or the string "gpu" to score the whole
    batch on the device in fp64 (bss_metrics.bss_eval_batch, hand-written kernels: no per-clip host loop, no D2H of the waveforms).
    Returns the reference's list
    [0, sdr, sir, sar, sdr_ls, sir_ls, si_sdr_ls, si_sdr].  `rng` seeds the all-zero-prediction guard
    (main.py:248-249 uses the unseeded global numpy RNG)."""
    batched = isinstance(get_metrics, str)
    if batched and get_metrics != "gpu":
        raise _lib.AvsepError(f"calc_metrics: unknown evaluator {get_metrics!r} (a callable or 'gpu')")
    if get_metrics is None:
        try:
            from asteroid.metrics import get_metrics as _gm   # noqa: the reference's own dependency
        except Exception as e:                                 # pragma: no cover
            raise _lib.AvsepError("calc_metrics: asteroid is not installed; pass get_metrics=...") from e
        get_metrics = _gm
    mag_mix = batch_data['mag_mix']
    phase_mix = batch_data['phase_mix']
    audios = batch_data['audios']
    pred_masks_ = outputs['pred_masks']
    N = args.num_mix
    B = mag_mix.size(0)
    dev = pred_masks_[0].device
    wav = ops.separate([pred_masks_[n] for n in range(N)], mag_mix.to(dev, non_blocking=True),
                       phase_mix.to(dev, non_blocking=True), log_freq=bool(args.log_freq),
                       binary_mask=bool(args.binary_mask), mask_thres=float(args.mask_thres), include_mix=True)
    if batched:
        preds = wav[:, 1:]                               # [B, N, L] view of the separated waveforms: scored in place
        L = preds.shape[-1]
        silent = (preds == 0).all(dim=-1).nonzero().tolist()
        if silent:                                       # main.py:248-249: all-zero prediction guard
            preds = preds.clone()
            for j, n in silent:
                r = rng.random(L) if rng is not None else np.random.rand(L)
                preds[j, n] = torch.from_numpy(0.01 * r).to(dev)
        gts = torch.stack([torch.as_tensor(audios[n])[:, 0:L].to(dev, non_blocking=True) for n in range(N)], dim=1)
        res = bss_metrics.bss_eval_batch(gts, preds)     # hand-written device evaluator: all B clips in one call
        per_clip = {k: res[k].mean(dim=1).tolist() for k in ('sdr', 'sir', 'sar', 'si_sdr')}      # asteroid average=True: mean over sources
        mean = lambda v: float(np.mean(v)) if v else 0.0
        return [0, mean(per_clip['sdr']), mean(per_clip['sir']), mean(per_clip['sar']), per_clip['sdr'], per_clip['sir'],
                per_clip['si_sdr'], mean(per_clip['si_sdr'])]
    wav = wav.double().cpu().numpy()                     # one D2H of [B, N+1, L] instead of N mask tensors
    sdr_m, sir_m, sar_m, sisdr_m = _Meter(), _Meter(), _Meter(), _Meter()
    sdr_ls, sir_ls, si_sdr_ls = [], [], []
    for j in range(B):
        mix_wav = wav[j, 0]
        preds_wav = [wav[j, 1 + n] for n in range(N)]
        for n in range(N):
            if np.all(preds_wav[n] == 0):
                r = rng.random(preds_wav[n].shape) if rng is not None else np.random.rand(*preds_wav[n].shape)
                preds_wav[n] = 0.01 * r
        L = preds_wav[0].shape[0]
        gts_wav = [np.asarray(audios[n][j, 0:L].cpu().numpy() if torch.is_tensor(audios[n]) else audios[n][j, 0:L])
                   for n in range(N)]
        md = get_metrics(mix_wav[0:L], np.asarray(gts_wav), np.asarray(preds_wav), sample_rate=11025,
                         metrics_list=['sdr', 'sir', 'sar', 'si_sdr'])
        sdr_m.update(md['sdr']); sir_m.update(md['sir']); sar_m.update(md['sar']); sisdr_m.update(md['si_sdr'])
        sdr_ls.append(md['sdr']); sir_ls.append(md['sir']); si_sdr_ls.append(md['si_sdr'])
    return [0, sdr_m.average(), sir_m.average(), sar_m.average(), sdr_ls, sir_ls, si_sdr_ls, sisdr_m.average()]


# ------------------------------------------------------------------------------------------------
# inference.py
# ------------------------------------------------------------------------------------------------
def get_spec(wav, device="cuda"):
    """inference.get_audfuncs()[1] = get_spec (inference.py:221-226): numpy wav [L] -> (amp, phase), each a CUDA tensor
    [1, 1, 512, T]; magnitude and angle come out of the STFT kernel's epilogue."""
    w = torch.as_tensor(np.ascontiguousarray(wav), dtype=torch.float32).to(device, non_blocking=True)
    amp, phase = ops.stft(w.reshape(1, -1), phase=True)
    return amp[:, None], phase[:, None]


_infer_cache = {}            # device index -> (mag_mix tensor, its version counter, warped magnitude, its log)


def prepare_inferdata(audios, frames, args, cache: bool = True):
    """NetWrapper.prepare_inferdata (inference.py:39-52): (mag_mix [B,1,512,T], phase_mix) -> (frames, mag_mix warped to the
    256-row log-frequency grid (+1e-10 before the warp), its log, phase_mix), in one fused kernel.

    The reference's vis_video calls this once per video frame on the SAME audio tensors (inference.py:537-569: 144-180 calls
    per clip), rebuilding the grid, the warp and the log every time.  With `cache` (default) a call whose `mag_mix` is the very
    tensor object of the previous call on that device, unmodified since (same autograd version counter), returns the cached
    warp / log without launching anything; the cached entry keeps `mag_mix` alive, so its
    storage cannot be recycled under the cache.  The returned tensors are shared between such calls: treat them as read-only,
    as the reference's callers do."""
    mag_mix, phase_mix = audios
    dev_key = mag_mix.device.index if torch.is_tensor(mag_mix) and mag_mix.is_cuda else None
    if cache and dev_key is not None:
        ent = _infer_cache.get(dev_key)
        if ent is not None and ent[0] is mag_mix and ent[1] == mag_mix._version:
            return frames, ent[2], ent[3], phase_mix
    _, mix_w, log_mix, _, _ = ops.prepare(mag_mix, [], log_freq=True, weighted_loss=False, binary_mask=True)
    if cache and dev_key is not None:
        _infer_cache[dev_key] = (mag_mix, mag_mix._version, mix_w, log_mix)
    return frames, mix_w, log_mix, phase_mix


def clear_inferdata_cache() -> None:
    """Drop the per-device prepare_inferdata cache (releases the cached spectrogram tensors)."""
    _infer_cache.clear()


def recover_aud_arrays(meta, audios, mask_thres: float = 0.5):
    """The arithmetic of inference.recover_aud (inference.py:433-491) on the GPU: un-warp the predicted masks to the
    linear frequency axis, threshold at 0.5 (the reference hard-codes `if True:` binary masks there), mask the mixture
    spectrogram, iSTFT.  meta['pred_masks']: list of N tensors [1,1,256,T]; audios = (mag_mix, phase_mix) [1,1,512,T].
    Returns dict(mix_wav [L'], preds_wav list of N [L'] (numpy float64, like istft_reconstruction), pred_masks list of
    N binary [256,T] numpy, pred_masks_linear list of N binary [512,T] numpy)."""
    pred_masks = meta['pred_masks']
    mag_mix, phase_mix = audios
    N = len(pred_masks)
    dev = pred_masks[0].device
    wav = ops.separate(list(pred_masks), mag_mix.to(dev), phase_mix.to(dev), log_freq=True, binary_mask=True,
                       mask_thres=mask_thres, include_mix=True)                       # [B, N+1, L']
    lin = [(ops.warp(pred_masks[n], ops.N_BINS, warp=False) > mask_thres).float() for n in range(N)]
    wav0 = wav[0].double().cpu().numpy()
    return dict(mix_wav=wav0[0], preds_wav=[wav0[1 + n] for n in range(N)],
                pred_masks=[(pred_masks[n][0, 0] > mask_thres).float().cpu().numpy() for n in range(N)],
                pred_masks_linear=[lin[n][0, 0].cpu().numpy() for n in range(N)])


def recover_aud(meta, audios, save_folder, write: Optional[Callable] = None, audRate: int = 11025):
    """inference.recover_aud (inference.py:433-491): recovers the mixture and the N predicted waveforms and hands them to
    `write(path, wav, rate)` (the reference calls soundfile.write(..., subtype='PCM_24'); soundfile / imageio are host
    I/O outside the hot path, so the writer is injected and the spectrogram / mask JPEGs are left to the caller, who gets
    the arrays from recover_aud_arrays).  Like the reference it rebinds meta['pred_masks'][n] to the binary numpy masks
    and returns the list of predicted-audio paths."""
    import os
    out = recover_aud_arrays(meta, audios)
    N = len(out['preds_wav'])
    if write is not None:
        write(os.path.join(save_folder, 'mix.wav'), out['mix_wav'], audRate)
    aud_pths = [None for _ in range(N)]
    for n in range(N):
        meta['pred_masks'][n] = out['pred_masks'][n][None, None]
        aud_pths[n] = os.path.join(save_folder, 'pred{}.wav'.format(n + 1))
        if write is not None:
            write(aud_pths[n], out['preds_wav'][n], audRate)
    return aud_pths


# ------------------------------------------------------------------------------------------------
# long-clip separation (BASELINE.json config C4)
# ------------------------------------------------------------------------------------------------
RECON_LEN = 65280      # samples a 65535-sample window reconstructs: hop * (T - 1) = 256 * 255


def clip_windows(audio: torch.Tensor, win: int = 65535, stride: int = RECON_LEN) -> torch.Tensor:
    """Cut a long clip [..., L] into windows [..., W, win]; the tail window is zero-padded, as the reference pads
    short clips (dataset/base.py:157-162).  The reference itself only ever processes one 6 s window
    (inference.py:197-228); this is the defined extension to 60 s clips.  The default stride is the length a window
    RECONSTRUCTS (65280 = 256 * (T - 1), not its 65535 input samples), so the reconstructed windows tile the clip
    without gaps."""
    L = audio.shape[-1]
    W = max(1, -(-max(L - win, 0) // stride) + 1)
    need = (W - 1) * stride + win
    if need > L:
        audio = torch.nn.functional.pad(audio, (0, need - L))
    return audio.unfold(-1, win, stride).contiguous()


def assemble_windows(wavs: torch.Tensor, stride: int, total_len: int) -> torch.Tensor:
    """Inverse of clip_windows for reconstructed windows [W, ..., Lw] -> [..., total_len]: window i supplies the
    samples [i*stride, i*stride + min(stride, Lw)), the last window everything it has; no cross-fade (the
    reference defines none).  With stride <= Lw (the default 65280) every sample up to the end of the last window is
    covered; a stride larger than Lw leaves the positions no window reconstructs at zero."""
    W, Lw = wavs.shape[0], wavs.shape[-1]
    out = wavs.new_zeros(wavs.shape[1:-1] + (total_len,))
    for i in range(W):
        n = Lw if i == W - 1 else min(stride, Lw)
        lo = i * stride
        hi = min(lo + n, total_len)
        if hi > lo:
            out[..., lo:hi] = wavs[i][..., :hi - lo]
    return out


def separate_long_clip(audio_mix: torch.Tensor, mask_fn: Callable, *, win: int = 65535, stride: int = RECON_LEN,
                       log_freq: bool = True, binary_mask: bool = True, mask_thres: float = 0.5,
                       rank: int = 0, world: int = 1) -> torch.Tensor:
    """Separate a long mixture [L] window by window on this rank's share of the windows.

    mask_fn(log_mag_mix [W,1,256,T], mag_mix_w [W,1,256,T]) -> masks [N, W, 256, T] (the model: U-Net + synthesizer).
    The STFT, warp and log of every window are computed once (the reference's vis_video recomputes them for each
    video frame, inference.py:557-563).  Returns this rank's reconstructed windows [W_local, N+1, 256 (T-1)];
    use parallel.gather_windows + assemble_windows to put the clip back together."""
    windows = clip_windows(audio_mix.reshape(-1), win, stride)
    mine = list(range(rank, windows.shape[0], world))
    w = windows[mine].contiguous()
    mag, _, re, im = ops.stft(w, phase=False, complex_parts=True)
    _, mix_w, log_mix, _, _ = ops.prepare(mag[:, None], [], log_freq=log_freq, weighted_loss=False, binary_mask=binary_mask)
    masks = mask_fn(log_mix, mix_w)
    return ops.separate(masks, None, None, re_mix=re, im_mix=im, log_freq=log_freq, binary_mask=binary_mask,
                        mask_thres=mask_thres, include_mix=True)


# ------------------------------------------------------------------------------------------------
# fused end-to-end steps (what bench.py measures)
# ------------------------------------------------------------------------------------------------
class _DevView:
    """Borrowed fp32 device memory exposed through __cuda_array_interface__."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr="<f4", data=(ptr, False), version=2)


class Step:
    """avsep_step: the whole hot path of one batch behind one C call (see include/avsep_b200.h)."""

    BUFFERS = dict(mag=0, mix_re=1, mix_im=2, mag_mix_w=3, log_mag_mix=4, weights=5, gt_masks=6, pred_masks=7,
                   loss=8, dvecs=9, dscale=10, dbias=11, mags_w=12, wav=13)

    def __init__(self, B: int, N: int, K: int = 32, wave_len: int = 65535):
        h = C.c_void_p()
        check(lib.avsep_step_create(C.byref(h), B, N, K, wave_len), "avsep_step_create")
        self._h = h
        self.B, self.N, self.K, self.L = B, N, K, wave_len
        self.T = 1 + wave_len // ops.HOP
        self.out_len = ops.HOP * (self.T - 1)

    @property
    def plan(self) -> "ops.Plan":
        """The step's STFT / iSTFT plan (borrowed): kernel-variant switches such as plan.set_mix_linear(False)."""
        return ops.Plan.borrowed(lib.avsep_step_plan(self._h))

    def set_graph(self, enable: bool = True):
        """Replay the train step as a CUDA graph (captured on the 2nd call with unchanged arguments; side streams only)."""
        check(lib.avsep_step_set_graph(self._h, int(bool(enable))), "avsep_step_set_graph")

    def set_fused_train(self, enable: bool = True):
        """False: run the separate synthesizer forward / backward kernels instead of the fused training kernel."""
        check(lib.avsep_step_set_fused_train(self._h, int(bool(enable))), "avsep_step_set_fused_train")

    def set_mix_on_device(self, enable: bool = True):
        """waves arguments become [N, B, L] (sources only); the mixture is summed on the GPU (dataset/music.py:127)."""
        check(lib.avsep_step_set_mix_on_device(self._h, int(bool(enable))), "avsep_step_set_mix_on_device")
        self.mix_on_device = bool(enable)

    def close(self):
        if getattr(self, "_h", None):
            torch.cuda.synchronize()
            lib.avsep_step_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, name):
        B, N, K, T = self.B, self.N, self.K, self.T
        return dict(mag=((N + 1) * B, 512, T), mix_re=(B, 512, T), mix_im=(B, 512, T), mag_mix_w=(B, 1, 256, T),
                    log_mag_mix=(B, 1, 256, T), weights=(B, 1, 256, T), gt_masks=(N, B, 1, 256, T),
                    pred_masks=(N, B, 1, 256, T), loss=(1,), dvecs=(B, N, K), dscale=(K,), dbias=(1,),
                    mags_w=(N, B, 1, 256, T), wav=(B, N + 1, self.out_len))[name]

    def buffer(self, name: str) -> torch.Tensor:
        """Copy of one of the step's device intermediates (tests / debugging)."""
        shape = self._shape(name)
        ptr = lib.avsep_step_buffer(self._h, self.BUFFERS[name])
        if not ptr:
            raise _lib.AvsepError(f"Step.buffer: no buffer named {name}")
        return torch.as_tensor(_DevView(int(ptr), shape), device="cuda").clone()

    def _check_waves(self, waves, what):
        """The C-ABI takes raw pointers: a [N, B, L] sources tensor handed to a step that expects [N+1, B, L] (or the
        reverse) would be read past its end, so the element count is checked here."""
        signals = self.N if getattr(self, "mix_on_device", False) else self.N + 1
        need = signals * self.B * self.L
        if waves.numel() != need or waves.dtype != torch.float32 or not waves.is_contiguous():
            raise _lib.AvsepError(f"{what}: waves must be contiguous float32 [{signals}, {self.B}, {self.L}] "
                                  f"({'sources only: the mixture is formed on the GPU' if signals == self.N else 'index 0 = mixture'}), "
                                  f"got shape {tuple(waves.shape)} {waves.dtype}")

    # waves: [N+1, B, L] (index 0 = mixture), or [N, B, L] sources after set_mix_on_device(True); vecs [B,N,K]; feat [B,K,256,T]
    def train_dev(self, waves, vecs, feat, scale, bias, dfeat, dvecs=None, dscale=None, dbias=None, loss=None):
        self._check_waves(waves, "Step.train_dev")
        check(lib.avsep_step_train_dev(self._h, waves.data_ptr(), vecs.data_ptr(), feat.data_ptr(), ops._p(scale), ops._p(bias),
                                       ops._p(dfeat), ops._p(dvecs), ops._p(dscale), ops._p(dbias), ops._p(loss), ops._stream()),
              "avsep_step_train_dev")

    def train_host(self, waves_host, vecs_host, feat, scale, bias, dfeat, dvecs_host=None):
        """Host buffers in (pinned CPU tensors), loss out as a python float; H2D / D2H inside the call."""
        self._check_waves(waves_host, "Step.train_host")
        loss = np.zeros(1, dtype=np.float32)
        check(lib.avsep_step_train_host(self._h, waves_host.data_ptr(), vecs_host.data_ptr(), feat.data_ptr(), ops._p(scale),
                                        ops._p(bias), ops._p(dfeat), ops._p(dvecs_host), loss.ctypes.data, ops._stream()),
              "avsep_step_train_host")
        return float(loss[0])

    def train_host_async(self, waves_host, vecs_host, feat, scale, bias, dfeat, loss_host, dvecs_host=None):
        """Enqueue H2D + step + D2H on the current stream and return; loss_host (pinned, 1 float) is valid after the
        stream is synchronised.  Two Steps on two streams overlap the copy of step i+1 with the kernels of step i."""
        self._check_waves(waves_host, "Step.train_host_async")
        check(lib.avsep_step_train_host_async(self._h, waves_host.data_ptr(), vecs_host.data_ptr(), feat.data_ptr(), ops._p(scale),
                                              ops._p(bias), ops._p(dfeat), ops._p(dvecs_host), loss_host.data_ptr(), ops._stream()),
              "avsep_step_train_host_async")

    def train_pcm16_async(self, pcm_host, gains_host, vecs_host, feat, scale, bias, dfeat, loss_host, dvecs_host=None):
        """Opt-in 16-bit PCM hand-off: pcm_host int16 [N+1,B,L] (or [N,B,L] sources only) pinned, gains_host float32
        [signals,B] or None; the loader's int16/32768 -> *gain -> clip -> /N runs on the device (dataset/base.py:156-172,
        dataset/music.py:120), bit-identical to the float32 hand-off at half the H2D bytes."""
        signals = self.N if getattr(self, "mix_on_device", False) else self.N + 1
        if pcm_host.dtype != torch.int16 or pcm_host.numel() != signals * self.B * self.L or not pcm_host.is_contiguous():
            raise _lib.AvsepError(f"Step.train_pcm16_async: pcm must be contiguous int16 [{signals}, {self.B}, {self.L}], "
                                  f"got {tuple(pcm_host.shape)} {pcm_host.dtype}")
        if gains_host is not None and (gains_host.dtype != torch.float32 or gains_host.numel() != signals * self.B):
            raise _lib.AvsepError(f"Step.train_pcm16_async: gains must be float32 [{signals}, {self.B}]")
        check(lib.avsep_step_train_pcm16_async(self._h, pcm_host.data_ptr(), ops._p(gains_host), vecs_host.data_ptr(), feat.data_ptr(),
                                               ops._p(scale), ops._p(bias), ops._p(dfeat), ops._p(dvecs_host), loss_host.data_ptr(),
                                               ops._stream()), "avsep_step_train_pcm16_async")

    def eval_host_async(self, waves_host, vecs_host, feat, scale, bias, wav_host=None, binary_mask=True, mask_thres=0.5):
        """Enqueue H2D + eval step (+ D2H of the waveforms when wav_host is given) on the current stream and return."""
        self._check_waves(waves_host, "Step.eval_host_async")
        check(lib.avsep_step_eval_host_async(self._h, waves_host.data_ptr(), vecs_host.data_ptr(), feat.data_ptr(), ops._p(scale),
                                             ops._p(bias), int(binary_mask), float(mask_thres), ops._p(wav_host), ops._stream()),
              "avsep_step_eval_host_async")

    def view(self, name: str) -> torch.Tensor:
        """Borrowed (no copy) view of one of the step's device buffers; valid until the step is closed."""
        ptr = lib.avsep_step_buffer(self._h, self.BUFFERS[name])
        if not ptr:
            raise _lib.AvsepError(f"Step.view: no buffer named {name}")
        return torch.as_tensor(_DevView(int(ptr), self._shape(name)), device="cuda")

    def eval_dev(self, waves, vecs, feat, scale, bias, wav=None, binary_mask=True, mask_thres=0.5):
        self._check_waves(waves, "Step.eval_dev")
        check(lib.avsep_step_eval_dev(self._h, waves.data_ptr(), vecs.data_ptr(), feat.data_ptr(), ops._p(scale), ops._p(bias),
                                      int(binary_mask), float(mask_thres), ops._p(wav), ops._stream()), "avsep_step_eval_dev")

    def eval_host(self, waves_host, vecs_host, feat, scale, bias, wav_host, binary_mask=True, mask_thres=0.5):
        self._check_waves(waves_host, "Step.eval_host")
        check(lib.avsep_step_eval_host(self._h, waves_host.data_ptr(), vecs_host.data_ptr(), feat.data_ptr(), ops._p(scale),
                                       ops._p(bias), int(binary_mask), float(mask_thres), wav_host.data_ptr(), ops._stream()),
              "avsep_step_eval_host")
        return wav_host


class StepPipeline:
    """`depth` Step objects on `depth` streams, used round-robin: the analogue of the reference's prefetching DataLoader
    (dataset/base.py:142-147 runs in worker processes while main.py:104-139 trains).  The STFT + prepare front end of a step
    depends on the waveforms only, so batch i + 1's STFT (tensor-bound) may run beside batch i's synthesizer kernels (HBM-bound);
    the persistent STFT grid is capped at `stft_sm_budget` SMs (avsep_plan_set_sm_budget) so the other step's CTAs find free SMs.
    Every step runs exactly the kernels of a lone Step (bit-identical results); measured +8..10 % train steps / s on a B200.

        pipe = StepPipeline(B, N, K)
        for i, batch in enumerate(loader):
            slot = pipe.next()                       # waits (stream-level) for nothing: the caller owns per-slot output buffers
            with torch.cuda.stream(slot.stream):
                slot.step.train_dev(batch.sources, vecs, feat, scale, bias, dfeat[slot.index], ..., loss[slot.index])
        pipe.synchronize()
    """

    class Slot:
        def __init__(self, index, step, stream):
            self.index, self.step, self.stream = index, step, stream

    def __init__(self, B: int, N: int, K: int = 32, wave_len: int = 65535, depth: int = 2, stft_sm_budget: int = 80,
                 mix_on_device: bool = True, graph: bool = True):
        if depth < 1:
            raise _lib.AvsepError(f"StepPipeline: depth must be >= 1 (got {depth})")
        self.slots = []
        for i in range(depth):
            st = Step(B, N, K, wave_len)
            st.set_mix_on_device(mix_on_device)
            st.set_graph(graph)
            if depth > 1 and stft_sm_budget:
                st.plan.set_sm_budget(stft_sm_budget)
            self.slots.append(StepPipeline.Slot(i, st, torch.cuda.Stream()))
        self._i = 0

    def next(self) -> "StepPipeline.Slot":
        """The slot of the next step; its stream first waits for the caller's current stream (inputs produced there)."""
        slot = self.slots[self._i % len(self.slots)]
        self._i += 1
        slot.stream.wait_stream(torch.cuda.current_stream())
        return slot

    def synchronize(self) -> None:
        """The caller's current stream waits for every slot (results consumed there afterwards)."""
        for slot in self.slots:
            torch.cuda.current_stream().wait_stream(slot.stream)

    def close(self) -> None:
        for slot in self.slots:
            slot.step.close()
        self.slots = []
