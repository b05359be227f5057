"""Tensor-level operators of the hot path: torch tensors in, torch tensors out, CUDA kernels inside.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); the arithmetic is in
csrc/*.cu behind the C-ABI.  Every op validates device / dtype / layout and raises RuntimeError
(AvsepError) like a torch op would -- there is no CPU path.
"""
from __future__ import annotations

import threading
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib, ptr_array

N_FFT = 1022
HOP = 256
N_BINS = 512
LOSS_KINDS = {"none": 0, "l1": 1, "l2": 2, "bce": 3}
ACTIVATIONS = {"no": 0, "none": 0, "sigmoid": 1}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.AvsepError(f"{name}: expected a CUDA tensor (avsep_b200 has no CPU path), got device {t.device}")
    if t.dtype != torch.float32:
        raise _lib.AvsepError(f"{name}: expected float32, got {t.dtype}")
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(lib.avsep_launch_count())


def device_check(device: Optional[int] = None) -> None:
    check(lib.avsep_device_check(-1 if device is None else int(device)), "avsep_device_check")


# ------------------------------------------------------------------------------------------------
# STFT / iSTFT plans
# ------------------------------------------------------------------------------------------------
class Plan:
    """Owner of an avsep_plan (DFT bases, TMA descriptors, bf16 staging buffers)."""

    def __init__(self, max_signals: int, wave_len: int):
        import ctypes as C
        h = C.c_void_p()
        check(lib.avsep_plan_create(C.byref(h), int(max_signals), int(wave_len)), "avsep_plan_create")
        self._h = h
        self.max_signals = int(max_signals)
        self.wave_len = int(wave_len)
        self.frames = int(lib.avsep_plan_frames(h))
        self.istft_len = int(lib.avsep_plan_istft_len(h))

    @classmethod
    def borrowed(cls, handle) -> "Plan":
        """Non-owning view of a plan that lives inside another object (avsep_step_plan); close() does nothing."""
        import ctypes as C
        self = cls.__new__(cls)
        self._h = C.c_void_p(handle) if isinstance(handle, int) else handle
        self._borrowed = True
        self.frames = int(lib.avsep_plan_frames(self._h))
        self.istft_len = int(lib.avsep_plan_istft_len(self._h))
        return self

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self) -> int:
        return int(lib.avsep_plan_workspace_bytes(self._h))

    def set_mix_linear(self, enable: bool) -> None:
        check(lib.avsep_plan_set_mix_linear(self._h, int(bool(enable))), "avsep_plan_set_mix_linear")

    def set_sm_budget(self, sms: int) -> None:
        """Persistent transform kernels of this plan use at most `sms` SMs (0 = all): SM-partitioned pipelines."""
        check(lib.avsep_plan_set_sm_budget(self._h, int(sms)), "avsep_plan_set_sm_budget")

    def set_spec_staged(self, enable: bool) -> None:
        """Inverse operand through the shared-memory staged kernel (default) or the register-tile kernel (cross-check)."""
        check(lib.avsep_plan_set_spec_staged(self._h, int(bool(enable))), "avsep_plan_set_spec_staged")

    def set_inverse_fold(self, enable: bool) -> None:
        check(lib.avsep_plan_set_inverse_fold(self._h, int(bool(enable))), "avsep_plan_set_inverse_fold")

    def set_fold(self, enable: bool) -> None:
        """False: forward transform through the plain 1024 x 1024 GEMM (the in-library cross-check of the folded kernels)."""
        check(lib.avsep_plan_set_fold(self._h, int(bool(enable))), "avsep_plan_set_fold")

    def close(self) -> None:
        if getattr(self, "_h", None) and not getattr(self, "_borrowed", False):
            lib.avsep_plan_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plans = {}
_plans_lock = threading.Lock()


def get_plan(num_signals: int, wave_len: int) -> Plan:
    """Plan cache keyed by (device, thread, stream, wave_len); capacity grows geometrically."""
    key = (torch.cuda.current_device(), threading.get_ident(), _stream(), int(wave_len))
    with _plans_lock:
        pl = _plans.get(key)
        if pl is None or pl.max_signals < num_signals:
            cap = max(int(num_signals), 1)
            if pl is not None:
                cap = max(cap, 2 * pl.max_signals)
                torch.cuda.synchronize()
                pl.close()
            pl = Plan(cap, wave_len)
            _plans[key] = pl
        return pl


def clear_plans() -> None:
    with _plans_lock:
        torch.cuda.synchronize()
        for pl in _plans.values():
            pl.close()
        _plans.clear()


def stft(wave: torch.Tensor, *, phase: bool = True, complex_parts: bool = False, plan: Optional[Plan] = None):
    """Batched librosa.stft(n_fft=1022, hop=256) -> (|.|, angle)  [dataset/base.py:142-147].

    wave [S, L] (or [L]) float32 CUDA -> mag [S,512,T] and, optionally, phase [S,512,T] and/or
    (re, im).  Returns (mag, phase | None) or (mag, phase | None, re, im) when complex_parts.
    """
    squeeze = wave.dim() == 1
    w = _cuda_f32(wave.reshape(1, -1) if squeeze else wave, "stft: wave")
    if w.dim() != 2:
        raise _lib.AvsepError(f"stft: wave must be [S, L] or [L], got shape {tuple(wave.shape)}")
    S, L = w.shape
    pl = plan or get_plan(S, L)
    T = pl.frames
    mag = torch.empty((S, N_BINS, T), device=w.device, dtype=torch.float32)
    ph = torch.empty_like(mag) if phase else None
    re = torch.empty_like(mag) if complex_parts else None
    im = torch.empty_like(mag) if complex_parts else None
    check(lib.avsep_stft(pl.handle, w.data_ptr(), S, mag.data_ptr(), _p(ph), _p(re), _p(im), _stream()), "avsep_stft")
    if squeeze:
        mag, ph = mag[0], (ph[0] if ph is not None else None)
        re, im = (re[0] if re is not None else None), (im[0] if im is not None else None)
    return (mag, ph, re, im) if complex_parts else (mag, ph)


def stft_mix(sources: torch.Tensor, *, phase: bool = False, plan: Optional[Plan] = None):
    """STFT magnitudes of B mixtures and their N sources from the sources alone [N, B, L]: returns
    (mag [(N+1), B, 512, T] with index 0 = the mixture = sum of the sources, phase | None).  The batched form of
    BaseDataset._mix_n_and_stft (dataset/base.py:174-189, mixture built at dataset/music.py:127); with N = 2 and no
    phase only the two source transforms run (the mixture spectrum is their sum)."""
    w = _cuda_f32(sources, "stft_mix: sources")
    if w.dim() != 3:
        raise _lib.AvsepError(f"stft_mix: sources must be [N, B, L], got shape {tuple(sources.shape)}")
    N, B, L = w.shape
    pl = plan or get_plan((N + 1) * B, L)
    T = pl.frames
    mag = torch.empty(((N + 1) * B, N_BINS, T), device=w.device, dtype=torch.float32)
    ph = torch.empty_like(mag) if phase else None
    check(lib.avsep_stft_mix(pl.handle, w.data_ptr(), B, N, mag.data_ptr(), _p(ph), _stream()), "avsep_stft_mix")
    shape = (N + 1, B, N_BINS, T)
    return mag.view(shape), (ph.view(shape) if ph is not None else None)


def pcm16_to_wave(pcm: torch.Tensor, gains: Optional[torch.Tensor] = None, num_mix: int = 1) -> torch.Tensor:
    """The loader's float path on the device: int16 [..., L] -> float32 = clip(int16 / 32768 * gain, -1, 1) / num_mix
    (dataset/base.py:156-172, dataset/music.py:120), bit-identical to numpy's float32 arithmetic."""
    if not pcm.is_cuda or pcm.dtype != torch.int16:
        raise _lib.AvsepError(f"pcm16_to_wave: expected a CUDA int16 tensor, got {pcm.dtype} on {pcm.device}")
    p_ = pcm.contiguous()
    L = p_.shape[-1]
    clips = p_.numel() // L if L else 0
    g = None
    if gains is not None:
        g = _cuda_f32(gains, "pcm16_to_wave: gains")
        if g.numel() != clips:
            raise _lib.AvsepError(f"pcm16_to_wave: {g.numel()} gains for {clips} clips")
    out = torch.empty(p_.shape, device=p_.device, dtype=torch.float32)
    check(lib.avsep_pcm16_to_wave(p_.data_ptr(), _p(g), clips, L, int(num_mix), out.data_ptr(), _stream()), "avsep_pcm16_to_wave")
    return out


def istft(mag: torch.Tensor, phase: torch.Tensor, *, plan: Optional[Plan] = None) -> torch.Tensor:
    """utils.istft_reconstruction (utils.py:101-104), batched: [S,512,T] x2 -> [S, 256 (T-1)] float32 in [-1, 1]."""
    m = _cuda_f32(mag, "istft: mag")
    p = _cuda_f32(phase, "istft: phase")
    if m.dim() != 3 or m.shape[1] != N_BINS or m.shape != p.shape:
        raise _lib.AvsepError(f"istft: expected mag/phase [S,512,T], got {tuple(mag.shape)} / {tuple(phase.shape)}")
    S, _, T = m.shape
    wave_len = HOP * (T - 1) + HOP - 1          # any length with 1 + L // 256 == T
    pl = plan or get_plan(S, wave_len)
    if pl.frames != T:
        raise _lib.AvsepError(f"istft: plan has {pl.frames} frames, input has {T}")
    out = torch.empty((S, pl.istft_len), device=m.device, dtype=torch.float32)
    check(lib.avsep_istft(pl.handle, m.data_ptr(), p.data_ptr(), S, out.data_ptr(), _stream()), "avsep_istft")
    return out


def separate(pred_masks, mag_mix: Optional[torch.Tensor], phase_mix: Optional[torch.Tensor], *,
             re_mix: Optional[torch.Tensor] = None, im_mix: Optional[torch.Tensor] = None,
             log_freq: bool = True, binary_mask: bool = True, mask_thres: float = 0.5,
             include_mix: bool = True, plan: Optional[Plan] = None) -> torch.Tensor:
    """Reconstruction half of calc_metrics (main.py:213-249) on the GPU.

    pred_masks: tensor [N,B,Hm,T] or list of N tensors [B,1,Hm,T] (Hm = 256 on the log grid).
    mixture: (mag_mix, phase_mix) [B,512,T] / [B,1,512,T], or (re_mix, im_mix).
    Returns wav [B, N + include_mix, 256 (T-1)]; index 0 is the mixture when include_mix.
    """
    if isinstance(pred_masks, (list, tuple)):
        masks = torch.stack([m.reshape(m.shape[0], m.shape[-2], m.shape[-1]) for m in pred_masks], dim=0)
    else:
        masks = pred_masks
    masks = _cuda_f32(masks, "separate: pred_masks")
    N, B, Hm, T = masks.shape
    if Hm != (256 if log_freq else N_BINS):
        raise _lib.AvsepError(f"separate: masks have {Hm} rows, expected {256 if log_freq else N_BINS}")

    def spec(t, name):
        if t is None:
            return None
        t = _cuda_f32(t.reshape(t.shape[0], t.shape[-2], t.shape[-1]), name)
        if tuple(t.shape) != (B, N_BINS, T):
            raise _lib.AvsepError(f"{name}: expected [{B},512,{T}], got {tuple(t.shape)}")
        return t

    mg, ph, re, im = spec(mag_mix, "separate: mag_mix"), spec(phase_mix, "separate: phase_mix"), \
        spec(re_mix, "separate: re_mix"), spec(im_mix, "separate: im_mix")
    nsig = N + (1 if include_mix else 0)
    wave_len = HOP * (T - 1) + HOP - 1
    pl = plan or get_plan(B * nsig, wave_len)
    out = torch.empty((B, nsig, pl.istft_len), device=masks.device, dtype=torch.float32)
    check(lib.avsep_separate(pl.handle, masks.data_ptr(), N, B, _p(mg), _p(ph), _p(re), _p(im), int(log_freq),
                             int(binary_mask), float(mask_thres), int(include_mix), out.data_ptr(), _stream()),
          "avsep_separate")
    return out


# ------------------------------------------------------------------------------------------------
# warp / prepare
# ------------------------------------------------------------------------------------------------
def warp(x: torch.Tensor, HO: int, warp: bool = True) -> torch.Tensor:
    """F.grid_sample(x, warpgrid(B, HO, T, warp), align_corners=False)  [main.py:62-66, 216-220].

    x [B,C,Hin,T] (or [S,Hin,T]) -> same rank with Hin replaced by HO.  No gradient (the reference
    only warps inputs that never require grad, main.py:52-66, and un-warps under no_grad, :423).
    """
    xs = _cuda_f32(x, "warp: x")
    shape = xs.shape
    if xs.dim() < 2:
        raise _lib.AvsepError(f"warp: expected [..., H, T], got {tuple(shape)}")
    Hin, T = shape[-2], shape[-1]
    S = xs.numel() // (Hin * T) if xs.numel() else 0
    y = torch.empty(shape[:-2] + (HO, T), device=xs.device, dtype=torch.float32)
    check(lib.avsep_warp(xs.data_ptr(), S, Hin, int(HO), T, int(bool(warp)), y.data_ptr(), _stream()), "avsep_warp")
    return y


def prepare(mag_mix: torch.Tensor, mags: Sequence[torch.Tensor], *, log_freq: bool = True, weighted_loss: bool = True,
            binary_mask: bool = True):
    """NetWrapper.prepare (main.py:51-95), tensor part, one fused kernel.

    mag_mix [B,1,512,T], mags: N x [B,1,512,T]  ->  (mags_w list, mag_mix_w, log_mag_mix, gt_masks list, weights),
    each [B,1,256,T] (or 512 rows when log_freq is off) -- the tuple NetWrapper.prepare returns.
    """
    mm = _cuda_f32(mag_mix, "prepare: mag_mix")
    if mm.dim() != 4 or mm.shape[1] != 1 or mm.shape[2] != N_BINS:
        raise _lib.AvsepError(f"prepare: mag_mix must be [B,1,512,T], got {tuple(mag_mix.shape)}")
    B, _, _, T = mm.shape
    N = len(mags)
    ms = [_cuda_f32(m, f"prepare: mags[{n}]") for n, m in enumerate(mags)]
    for n, m in enumerate(ms):
        if m.shape != mm.shape:
            raise _lib.AvsepError(f"prepare: mags[{n}] shape {tuple(m.shape)} != mag_mix shape {tuple(mm.shape)}")
    HO = 256 if log_freq else N_BINS
    new = lambda: torch.empty((B, 1, HO, T), device=mm.device, dtype=torch.float32)
    mix_w, log_mix, weights = new(), new(), new()
    mags_w = [new() for _ in range(N)]
    gts = [new() for _ in range(N)]
    check(lib.avsep_prepare(mm.data_ptr(), ptr_array([m.data_ptr() for m in ms]), N, B, T, int(log_freq),
                            int(weighted_loss), int(binary_mask), mix_w.data_ptr(),
                            ptr_array([m.data_ptr() for m in mags_w]), log_mix.data_ptr(), weights.data_ptr(),
                            ptr_array([g.data_ptr() for g in gts]), _stream()), "avsep_prepare")
    return mags_w, mix_w, log_mix, gts, weights


# ------------------------------------------------------------------------------------------------
# synthesizer (+ sigmoid, + loss) with autograd
# ------------------------------------------------------------------------------------------------
def _scratch(B, N, K, HW, device) -> torch.Tensor:
    n = int(lib.avsep_synth_scratch_bytes(B, N, K, HW))
    return torch.empty(n, dtype=torch.uint8, device=device)


def _synth_forward_raw(feat, vecs, scale, bias, activation, loss_kind=0, gt=None, weight=None):
    B, K = feat.shape[0], feat.shape[1]
    HW = feat[0, 0].numel()
    N = vecs.shape[1]
    out = torch.empty((N, B) + tuple(feat.shape[2:]), device=feat.device, dtype=torch.float32)
    loss = torch.zeros((), device=feat.device, dtype=torch.float32) if loss_kind else None
    scratch = _scratch(B, N, K, HW, feat.device)
    check(lib.avsep_synth_forward(feat.data_ptr(), vecs.data_ptr(), _p(scale), _p(bias), B, N, K, HW, activation,
                                  out.data_ptr(), loss_kind, _p(gt), _p(weight), _p(loss), scratch.data_ptr(), _stream()),
          "avsep_synth_forward")
    return out, loss


class _SynthMasks(torch.autograd.Function):
    """out[n] = act(sum_k vecs[:, n, k] * scale[k] * feat[:, k] + bias) for all n in one pass."""

    @staticmethod
    def forward(ctx, feat, vecs, scale, bias, activation):
        f = _cuda_f32(feat, "synth: feat_sound")
        v = _cuda_f32(vecs, "synth: vecs")
        s = _cuda_f32(scale, "synth: scale") if scale is not None else None
        b = _cuda_f32(bias, "synth: bias") if bias is not None else None
        out, _ = _synth_forward_raw(f, v, s, b, activation)
        ctx.save_for_backward(f, v, s if s is not None else torch.empty(0, device=f.device), out)
        ctx.has_scale = s is not None
        ctx.has_bias = b is not None
        ctx.activation = activation
        return out

    @staticmethod
    def backward(ctx, gout):
        f, v, s, out = ctx.saved_tensors
        s = s if ctx.has_scale else None
        B, K = f.shape[0], f.shape[1]
        HW = f[0, 0].numel()
        N = v.shape[1]
        g = _cuda_f32(gout, "synth backward: grad")
        if ctx.activation == 1:                      # sigmoid': p (1 - p), elementwise on the small [N,B,HW] tensor
            g = g * out * (1.0 - out)
        need = ctx.needs_input_grad
        dfeat = torch.empty_like(f) if need[0] else None
        dvecs = torch.empty_like(v) if need[1] else None
        dscale = torch.empty_like(s) if (s is not None and need[2]) else None
        dbias = torch.empty(1, device=f.device, dtype=torch.float32) if (ctx.has_bias and need[3]) else None
        scratch = _scratch(B, N, K, HW, f.device)
        check(lib.avsep_synth_backward(f.data_ptr(), v.data_ptr(), _p(s), B, N, K, HW, 0, 0, None, None, None, None,
                                       g.contiguous().data_ptr(), _p(dfeat), _p(dvecs), _p(dscale), _p(dbias),
                                       scratch.data_ptr(), _stream()), "avsep_synth_backward")
        return dfeat, dvecs, dscale, dbias, None


class _SynthMasksLoss(torch.autograd.Function):
    """Fused synthesizer + activation + list-form weighted loss.

    When feat_sound needs a gradient (and K <= 32) the forward call runs the fused training kernel: loss AND the
    gradients for an upstream gradient of 1 come out of a single pass over feat_sound; backward only rescales them
    by the actual upstream scalar (a no-op kernel when it is 1).  Otherwise forward saves the masks and backward
    recomputes dz from them (second pass over feat_sound)."""

    @staticmethod
    def forward(ctx, feat, vecs, scale, bias, gt, weight, activation, loss_kind, allow_fused=True):
        f = _cuda_f32(feat, "synth_loss: feat_sound")
        v = _cuda_f32(vecs, "synth_loss: vecs")
        s = _cuda_f32(scale, "synth_loss: scale") if scale is not None else None
        b = _cuda_f32(bias, "synth_loss: bias") if bias is not None else None
        g = _cuda_f32(gt, "synth_loss: gt_masks")
        w = _cuda_f32(weight, "synth_loss: weight") if weight is not None else None
        B, K = f.shape[0], f.shape[1]
        HW = f[0, 0].numel()
        N = v.shape[1]
        empty = torch.empty(0, device=f.device)
        ctx.has_scale, ctx.has_bias, ctx.has_weight = s is not None, b is not None, w is not None
        ctx.activation, ctx.loss_kind = activation, loss_kind
        ctx.fused = bool(allow_fused) and bool(ctx.needs_input_grad[0]) and K <= 32 and HW % 4 == 0
        ctx.consumed = False
        if ctx.fused:
            out = torch.empty((N, B) + tuple(f.shape[2:]), device=f.device, dtype=torch.float32)
            loss = torch.zeros((), device=f.device, dtype=torch.float32)
            dfeat = torch.empty_like(f)
            dvecs = torch.empty_like(v)
            dscale = torch.empty_like(s) if s is not None else None
            dbias = torch.empty(1, device=f.device, dtype=torch.float32)
            scratch = _scratch(B, N, K, HW, f.device)
            check(lib.avsep_synth_train(f.data_ptr(), v.data_ptr(), _p(s), _p(b), B, N, K, HW, activation, loss_kind,
                                        g.data_ptr(), _p(w), None, out.data_ptr(), loss.data_ptr(), dfeat.data_ptr(),
                                        dvecs.data_ptr(), _p(dscale), dbias.data_ptr(), scratch.data_ptr(), _stream()),
                  "avsep_synth_train")
            ctx.save_for_backward(dfeat, dvecs, dscale if dscale is not None else empty, dbias)
        else:
            out, loss = _synth_forward_raw(f, v, s, b, activation, loss_kind, g, w)
            ctx.save_for_backward(f, v, s if s is not None else empty, out, g, w if w is not None else empty)
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, gloss, _gout):
        gs = _cuda_f32(gloss.reshape(1), "synth_loss backward: grad")
        need = ctx.needs_input_grad
        if ctx.fused:
            # The saved buffers ARE the gradients (for an upstream gradient of 1) and are rescaled in place below, then
            # handed to autograd, which may keep them as .grad without a copy.  A second backward over the same graph
            # (retain_graph=True, a gradient penalty, two losses sharing err) would rescale them again -- g1 * g2 -- and
            # alias what the first call returned, so it is refused instead of returning wrong gradients silently.
            if ctx.consumed:
                raise RuntimeError("synth_masks_loss: the fused training kernel's gradients were already consumed by an earlier "
                                   "backward over this graph; call synth_masks_loss(..., fused=False) (separate, re-entrant "
                                   "forward / backward kernels) when the graph is backpropagated more than once")
            ctx.consumed = True
            dfeat, dvecs, dscale, dbias = ctx.saved_tensors
            outs = [dfeat, dvecs if need[1] else None, dscale if (ctx.has_scale and need[2]) else None,
                    dbias if (ctx.has_bias and need[3]) else None]
            for t in outs:
                if t is not None:
                    check(lib.avsep_scale_if_needed(t.data_ptr(), t.numel(), gs.data_ptr(), _stream()), "avsep_scale_if_needed")
            return outs[0], outs[1], outs[2], outs[3], None, None, None, None, None
        f, v, s, out, g, w = ctx.saved_tensors
        s = s if ctx.has_scale else None
        w = w if ctx.has_weight else None
        B, K = f.shape[0], f.shape[1]
        HW = f[0, 0].numel()
        N = v.shape[1]
        dfeat = torch.empty_like(f) if need[0] else None
        dvecs = torch.empty_like(v) if need[1] else None
        dscale = torch.empty_like(s) if (s is not None and need[2]) else None
        dbias = torch.empty(1, device=f.device, dtype=torch.float32) if (ctx.has_bias and need[3]) else None
        scratch = _scratch(B, N, K, HW, f.device)
        check(lib.avsep_synth_backward(f.data_ptr(), v.data_ptr(), _p(s), B, N, K, HW, ctx.activation, ctx.loss_kind,
                                       out.data_ptr(), g.data_ptr(), _p(w), gs.data_ptr(), None, _p(dfeat), _p(dvecs),
                                       _p(dscale), _p(dbias), scratch.data_ptr(), _stream()), "avsep_synth_backward")
        return dfeat, dvecs, dscale, dbias, None, None, None, None, None


def _check_synth_shapes(feat, vecs, scale):
    if feat.dim() < 3:
        raise _lib.AvsepError(f"synth: feat_sound must be [B,K,...], got {tuple(feat.shape)}")
    if vecs.dim() != 3 or vecs.shape[0] != feat.shape[0] or vecs.shape[2] != feat.shape[1]:
        raise _lib.AvsepError(f"synth: vecs must be [B,N,K] = [{feat.shape[0]},N,{feat.shape[1]}], got {tuple(vecs.shape)}")
    if not 1 <= vecs.shape[1] <= 4:
        raise _lib.AvsepError(f"synth: N = {vecs.shape[1]} sources per call outside [1, 4]")
    if scale is not None and tuple(scale.shape) != (feat.shape[1],):
        raise _lib.AvsepError(f"synth: scale must be [{feat.shape[1]}], got {tuple(scale.shape)}")


def synth_masks(feat_sound: torch.Tensor, vecs: torch.Tensor, scale: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                activation: str = "sigmoid") -> torch.Tensor:
    """All N synthesizer outputs in one pass: returns [N, B, *spatial]; differentiable."""
    _check_synth_shapes(feat_sound, vecs, scale)
    if activation not in ACTIVATIONS:
        raise Exception("Unkown activation!")          # models/__init__.py:28
    return _SynthMasks.apply(feat_sound, vecs, scale, bias, ACTIVATIONS[activation])


def synth_masks_loss(feat_sound, vecs, scale, bias, gt_masks, weight, *, loss: str = "bce", activation: str = "sigmoid",
                     fused: bool = True):
    """Fused SoP++ step (SoP++/main.py:203-209): returns (err scalar, masks [N,B,*spatial] detached).

    fused=True (default) lets the forward call run the one-pass training kernel when feat_sound needs a gradient; its
    graph can be backpropagated ONCE (a second backward raises).  fused=False always uses the separate forward /
    backward kernels, which read feat_sound twice but may be backpropagated any number of times (retain_graph)."""
    _check_synth_shapes(feat_sound, vecs, scale)
    if isinstance(gt_masks, (list, tuple)):
        gt_masks = torch.stack([g.reshape(g.shape[0], *feat_sound.shape[2:]) for g in gt_masks], dim=0)
    N, B = vecs.shape[1], feat_sound.shape[0]
    if gt_masks.numel() != N * B * feat_sound[0, 0].numel():
        raise _lib.AvsepError(f"synth_masks_loss: gt_masks has {gt_masks.numel()} elements, expected [N,B,*spatial]")
    if weight is not None and weight.numel() != B * feat_sound[0, 0].numel():
        raise _lib.AvsepError("synth_masks_loss: weight must be [B,1,*spatial]")
    if loss not in ("bce", "l1", "l2"):
        raise Exception("Architecture undefined!")     # models/__init__.py:129
    return _SynthMasksLoss.apply(feat_sound, vecs, scale, bias, gt_masks, weight, ACTIVATIONS[activation], LOSS_KINDS[loss],
                                 bool(fused))


def synth_nosum(feat_img, feat_sound, scale, bias) -> torch.Tensor:
    """InnerProd.forward_nosum (synthesizer_net.py:21-26), inference helper (no autograd)."""
    f = _cuda_f32(feat_sound, "forward_nosum: feat_sound")
    B, K = f.shape[0], f.shape[1]
    v = _cuda_f32(feat_img.reshape(B, K), "forward_nosum: feat_img")
    out = torch.empty_like(f)
    check(lib.avsep_synth_nosum(f.data_ptr(), v.data_ptr(), _p(_cuda_f32(scale, "scale") if scale is not None else None),
                                _p(_cuda_f32(bias, "bias") if bias is not None else None), B, K, f[0, 0].numel(),
                                out.data_ptr(), _stream()), "avsep_synth_nosum")
    return out


def synth_pixelwise(feats_img, feat_sound, scale, bias) -> torch.Tensor:
    """InnerProd.forward_pixelwise (synthesizer_net.py:29-38): [B,K,HI,WI] x [B,K,HS,WS] -> [B,HI,WI,HS,WS]."""
    f = _cuda_f32(feat_sound, "forward_pixelwise: feat_sound")
    fi = _cuda_f32(feats_img, "forward_pixelwise: feats_img")
    B, K, HI, WI = fi.shape
    _, _, HS, WS = f.shape
    out = torch.empty((B, HI, WI, HS, WS), device=f.device, dtype=torch.float32)
    check(lib.avsep_synth_pixelwise(f.data_ptr(), fi.data_ptr(), _p(_cuda_f32(scale, "scale") if scale is not None else None),
                                    _p(_cuda_f32(bias, "bias") if bias is not None else None), B, K, HI * WI, HS * WS,
                                    out.data_ptr(), _stream()), "avsep_synth_pixelwise")
    return out


# ------------------------------------------------------------------------------------------------
# criterion
# ------------------------------------------------------------------------------------------------
class _ListLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, weight, n, *tensors):
        preds = [_cuda_f32(t, f"loss: preds[{i}]") for i, t in enumerate(tensors[:n])]
        tgts = [_cuda_f32(t, f"loss: targets[{i}]") for i, t in enumerate(tensors[n:])]
        numel = preds[0].numel()
        for i in range(n):
            if preds[i].numel() != numel or tgts[i].numel() != numel:
                raise _lib.AvsepError("loss: all predictions and targets must have the same number of elements")
        w = None
        if weight is not None:
            w = _cuda_f32(weight, "loss: weight")
            if w.numel() != numel:
                w = w.expand_as(preds[0]).contiguous()
        loss = torch.zeros((), device=preds[0].device, dtype=torch.float32)
        scratch = torch.empty(int(lib.avsep_loss_scratch_bytes(n, numel)), dtype=torch.uint8, device=preds[0].device)
        check(lib.avsep_loss_forward(kind, ptr_array([p.data_ptr() for p in preds]), ptr_array([t.data_ptr() for t in tgts]),
                                     _p(w), n, numel, loss.data_ptr(), scratch.data_ptr(), _stream()), "avsep_loss_forward")
        ctx.kind, ctx.n, ctx.has_w = kind, n, w is not None
        ctx.save_for_backward(*(preds + tgts + ([w] if w is not None else [])))
        return loss

    @staticmethod
    def backward(ctx, gloss):
        saved = ctx.saved_tensors
        n = ctx.n
        preds, tgts = list(saved[:n]), list(saved[n:2 * n])
        w = saved[2 * n] if ctx.has_w else None
        gs = _cuda_f32(gloss.reshape(1), "loss backward: grad")
        grads = [torch.empty_like(p) if ctx.needs_input_grad[3 + i] else None for i, p in enumerate(preds)]
        check(lib.avsep_loss_backward(ctx.kind, ptr_array([p.data_ptr() for p in preds]),
                                      ptr_array([t.data_ptr() for t in tgts]), _p(w), n, preds[0].numel(), gs.data_ptr(),
                                      ptr_array([_p(g) for g in grads]), _stream()), "avsep_loss_backward")
        return (None, None, None, *grads, *([None] * n))


def list_loss(kind: str, preds: Sequence[torch.Tensor], targets: Sequence[torch.Tensor], weight: Optional[torch.Tensor]):
    """BaseLoss.forward list form (criterion.py:10-25): mean_n mean(w * l(pred_n, tgt_n)); differentiable in preds."""
    n = len(preds)
    if n < 1 or n > 8 or len(targets) != n:
        raise _lib.AvsepError(f"loss: need 1..8 predictions with as many targets, got {n} / {len(targets)}")
    return _ListLoss.apply(LOSS_KINDS[kind], weight, n, *preds, *targets)


def pit_lossmat(kind: str, preds: torch.Tensor, targets: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Per-sample S x S loss matrix of PitWrapper._opt_perm_loss (criterion.py:159-176): [B, S, S]."""
    p = _cuda_f32(preds, "pit: preds")
    t = _cuda_f32(targets, "pit: targets")
    w = _cuda_f32(weights, "pit: weights")
    if p.shape != t.shape or p.shape != w.shape or p.dim() < 2:
        raise _lib.AvsepError("pit: preds, targets and weights must share the shape [B, ..., S]")
    B, S = p.shape[0], p.shape[-1]
    P = p[0].numel() // S
    out = torch.empty((B, S, S), device=p.device, dtype=torch.float32)
    check(lib.avsep_pit_lossmat(LOSS_KINDS[kind], p.data_ptr(), t.data_ptr(), w.data_ptr(), B, P, S, out.data_ptr(), _stream()),
          "avsep_pit_lossmat")
    return out


class Perms:
    """The optimal permutations of a batch, kept on the device: `tensor` is int32 [B, S].  Behaves like the reference's
    `perms` (a list of tuples: len(), indexing, iteration, ==) but only copies to the host -- one synchronising D2H of
    B * S integers -- when Python actually looks at the values; PitWrapper.reorder_tensor consumes it on the device."""

    def __init__(self, tensor: torch.Tensor, inverse: torch.Tensor):
        self.tensor = tensor
        self.inverse = inverse
        self._host = None

    def tolist(self):
        if self._host is None:
            self._host = [tuple(p) for p in self.tensor.tolist()]
        return self._host

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, i):
        return self.tolist()[i]

    def __iter__(self):
        return iter(self.tolist())

    def __eq__(self, other):
        return self.tolist() == [tuple(p) for p in other]

    def __repr__(self):
        return repr(self.tolist())


def pit_search(lossmat: torch.Tensor):
    """PitWrapper._fast_pit (criterion.py:113-136) for the whole batch on the device: [B,S,S] -> (loss [B], Perms)."""
    lm = _cuda_f32(lossmat, "pit_search: lossmat")
    B, S = lm.shape[0], lm.shape[1]
    loss = torch.empty(B, device=lm.device, dtype=torch.float32)
    perm = torch.empty((B, S), device=lm.device, dtype=torch.int32)
    inv = torch.empty((B, S), device=lm.device, dtype=torch.int32)
    check(lib.avsep_pit_search(lm.data_ptr(), B, S, loss.data_ptr(), perm.data_ptr(), inv.data_ptr(), _stream()), "avsep_pit_search")
    return loss, Perms(perm, inv)


def pit_reorder(tensor: torch.Tensor, perms) -> torch.Tensor:
    """PitWrapper.reorder_tensor (criterion.py:180-200): out[b, ..., s] = tensor[b, ..., perms[b][s]]."""
    x = _cuda_f32(tensor, "pit_reorder: tensor")
    B, S = x.shape[0], x.shape[-1]
    pt = perms.tensor if isinstance(perms, Perms) else torch.as_tensor([list(p) for p in perms], dtype=torch.int32, device=x.device)
    if tuple(pt.shape) != (B, S):
        raise _lib.AvsepError(f"pit_reorder: {tuple(pt.shape)} permutations for a tensor of shape {tuple(x.shape)}")
    y = torch.empty_like(x)
    check(lib.avsep_pit_reorder(x.data_ptr(), pt.contiguous().data_ptr(), B, x[0].numel() // S, S, y.data_ptr(), _stream()), "avsep_pit_reorder")
    return y


class _PitLoss(torch.autograd.Function):
    """Carries the gradient of loss[b] = mean_i lossmat[b][i][perm_b[i]] back to preds (one streaming kernel)."""

    @staticmethod
    def forward(ctx, kind, preds, targets, weights, inv_perm, loss_values):
        ctx.kind = kind
        ctx.save_for_backward(preds, targets, weights, inv_perm)
        return loss_values.clone()

    @staticmethod
    def backward(ctx, gloss):
        p, t, w, inv = ctx.saved_tensors
        B, S = p.shape[0], p.shape[-1]
        P = p[0].numel() // S
        g = _cuda_f32(gloss.reshape(B), "pit backward: grad")
        dp = torch.empty_like(p)
        check(lib.avsep_pit_backward(LOSS_KINDS[ctx.kind], p.data_ptr(), t.data_ptr(), w.data_ptr(), inv.data_ptr(), g.data_ptr(),
                                     B, P, S, dp.data_ptr(), _stream()), "avsep_pit_backward")
        return None, dp, None, None, None, None


def _plane_gather(x: torch.Tensor, sel: Optional[torch.Tensor], rowscale: Optional[torch.Tensor] = None, factor: float = 1.0) -> torch.Tensor:
    """y[j][b] = x[sel[b][j]][b] * (rowscale[b] * factor) for x [N, B, ...] (avsep_plane_gather)."""
    N, B = x.shape[0], x.shape[1]
    y = torch.empty_like(x)
    check(lib.avsep_plane_gather(x.data_ptr(), _p(sel), _p(rowscale), float(factor), N, B, x[0, 0].numel(), y.data_ptr(), _stream()),
          "avsep_plane_gather")
    return y


class _SynthPitLoss(torch.autograd.Function):
    """SoP++ audio-only step (SoP++/main.py:227-238): N synthesizer masks + sigmoid + per-sample PIT loss.

    forward: ONE pass over feat_sound gives the masks and the [B,N,N] loss matrix (avsep_synth_forward_pit), the
    permutation search runs on the device (avsep_pit_search).  backward: with the permutation fixed, the loss is the
    list-form weighted loss against per-sample re-ordered targets, so the synthesizer's own backward kernel is reused
    (targets gathered by the inverse permutation, per-sample upstream gradients folded into the weights)."""

    @staticmethod
    def forward(ctx, feat, vecs, scale, bias, gt, weight, activation, loss_kind):
        f = _cuda_f32(feat, "synth_pit_loss: feat_sound")
        v = _cuda_f32(vecs, "synth_pit_loss: vecs")
        s = _cuda_f32(scale, "synth_pit_loss: scale") if scale is not None else None
        b = _cuda_f32(bias, "synth_pit_loss: bias") if bias is not None else None
        g = _cuda_f32(gt, "synth_pit_loss: gt_masks")
        B, K = f.shape[0], f.shape[1]
        HW = f[0, 0].numel()
        N = v.shape[1]
        w = _cuda_f32(weight, "synth_pit_loss: weight") if weight is not None else torch.ones((B, HW), device=f.device)
        out = torch.empty((N, B) + tuple(f.shape[2:]), device=f.device, dtype=torch.float32)
        lossmat = torch.empty((B, N, N), device=f.device, dtype=torch.float32)
        scratch = _scratch(B, N, K, HW, f.device)
        check(lib.avsep_synth_forward_pit(f.data_ptr(), v.data_ptr(), _p(s), _p(b), B, N, K, HW, activation, out.data_ptr(), loss_kind,
                                          g.data_ptr(), w.data_ptr(), lossmat.data_ptr(), scratch.data_ptr(), _stream()),
              "avsep_synth_forward_pit")
        loss, perms = pit_search(lossmat)
        ordered = _plane_gather(out, perms.tensor)                   # ordered[i][b] = out[perm[b][i]][b]  (reorder_tensor)
        empty = torch.empty(0, device=f.device)
        ctx.save_for_backward(f, v, s if s is not None else empty, out, g, w, perms.inverse)
        ctx.has_scale, ctx.has_bias = s is not None, b is not None
        ctx.activation, ctx.loss_kind = activation, loss_kind
        ctx.perms = perms
        ctx.mark_non_differentiable(ordered, lossmat)
        return loss, ordered, lossmat

    @staticmethod
    def backward(ctx, gloss, _gord, _glm):
        f, v, s, out, g, w, inv = ctx.saved_tensors
        s = s if ctx.has_scale else None
        B, K = f.shape[0], f.shape[1]
        HW = f[0, 0].numel()
        N = v.shape[1]
        need = ctx.needs_input_grad
        gl = _cuda_f32(gloss.reshape(B), "synth_pit_loss backward: grad")
        # loss[b] = 1/(N HW) sum_j sum_px w l(pred_j, gt[inv[b][j]])  =>  sum_b gloss[b] loss[b] is the list-form mean loss over
        # (N, B, HW) with targets re-ordered per sample and weights scaled by gloss[b] * B
        gt_perm = _plane_gather(g.reshape(N, B, HW), inv)
        w_scaled = _plane_gather(w.reshape(1, B, HW), None, gl, float(B)).reshape(B, HW)
        one = torch.ones(1, device=f.device)
        dfeat = torch.empty_like(f) if need[0] else None
        dvecs = torch.empty_like(v) if need[1] else None
        dscale = torch.empty_like(s) if (s is not None and need[2]) else None
        dbias = torch.empty(1, device=f.device, dtype=torch.float32) if (ctx.has_bias and need[3]) else None
        scratch = _scratch(B, N, K, HW, f.device)
        check(lib.avsep_synth_backward(f.data_ptr(), v.data_ptr(), _p(s), B, N, K, HW, ctx.activation, ctx.loss_kind,
                                       out.data_ptr(), gt_perm.data_ptr(), w_scaled.data_ptr(), one.data_ptr(), None, _p(dfeat), _p(dvecs),
                                       _p(dscale), _p(dbias), scratch.data_ptr(), _stream()), "avsep_synth_backward")
        return dfeat, dvecs, dscale, dbias, None, None, None, None


def synth_masks_pit_loss(feat_sound, vecs, scale, bias, gt_masks, weight, *, loss: str = "bce", activation: str = "sigmoid"):
    """SoP++ audio-only step (SoP++/main.py:227-238), fused: returns (err [B] differentiable, perms (device-backed, see Perms),
    ordered masks [N,B,*spatial] detached = reorder_tensor(pred_masks, perms), lossmat [B,N,N])."""
    _check_synth_shapes(feat_sound, vecs, scale)
    if isinstance(gt_masks, (list, tuple)):
        gt_masks = torch.stack([g.reshape(g.shape[0], *feat_sound.shape[2:]) for g in gt_masks], dim=0)
    N, B = vecs.shape[1], feat_sound.shape[0]
    if gt_masks.numel() != N * B * feat_sound[0, 0].numel():
        raise _lib.AvsepError(f"synth_masks_pit_loss: gt_masks has {gt_masks.numel()} elements, expected [N,B,*spatial]")
    if feat_sound[0, 0].numel() % 4 != 0:
        raise _lib.AvsepError("synth_masks_pit_loss: H * W must be a multiple of 4")
    if loss not in ("bce", "l1", "l2"):
        raise Exception("Architecture undefined!")     # models/__init__.py:129
    err, ordered, lossmat = _SynthPitLoss.apply(feat_sound, vecs, scale, bias, gt_masks, weight, ACTIVATIONS[activation], LOSS_KINDS[loss])
    perms = err.grad_fn.perms if err.grad_fn is not None and hasattr(err.grad_fn, "perms") else pit_search(lossmat)[1]
    return err, perms, ordered, lossmat


def pit_loss(kind: str, preds: torch.Tensor, targets: torch.Tensor, weights: torch.Tensor):
    """PitWrapper.forward (criterion.py:202-231): returns (loss [B], differentiable in preds; perms list of tuples).

    The S x S loss matrix of every sample comes from one kernel, the S! search from another: nothing is copied to the
    host and the stream is never synchronised; `perms` is a device-backed sequence (ops.Perms)."""
    p = _cuda_f32(preds, "pit: preds")
    t = _cuda_f32(targets, "pit: targets")
    w = _cuda_f32(weights, "pit: weights")
    lm = pit_lossmat(kind, p, t, w)
    loss_t, perms = pit_search(lm)
    inv_t = perms.inverse
    if preds.requires_grad and torch.is_grad_enabled():
        if preds.data_ptr() != p.data_ptr():                         # preds was copied (non-contiguous / unaligned): keep the graph
            return _PitLoss.apply(kind, preds.contiguous(), t, w, inv_t, loss_t), perms
        return _PitLoss.apply(kind, preds, t, w, inv_t, loss_t), perms
    return loss_t, perms
