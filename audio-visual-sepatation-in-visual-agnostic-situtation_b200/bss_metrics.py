"""Batched SDR / SIR / SAR / SI-SDR on the device.

The reference scores every clip on the host, one at a time, through
``asteroid.metrics.get_metrics(mix, clean, estimate, sample_rate=11025,
metrics_list=['sdr','sir','sar','si_sdr'])`` (main.py:252-275), which routes to
``mir_eval.separation.bss_eval_sources(compute_permutation=False)`` (BSS-eval v3, 512-tap
distortion filters) and ``pb_bss_eval.evaluation.si_sdr``.  Once the iSTFT runs on the GPU that loop
is the whole evaluation wall time, so all B clips of a batch are scored at once, in fp64, where the
waveforms already are.

* ``bss_eval_batch``  -- the product path: hand-written kernels behind ``avsep_bss_eval`` (csrc/bss_eval.cu): direct
  lag-0..511 cross-correlations, block-Toeplitz Levinson solves of the Gram systems, projection energies as quadratic
  forms.  CUDA tensors only (there is no CPU path), float32 or float64, strided views are scored in place.
* ``bss_eval_batch_torch`` -- the round-1 implementation through ``torch.fft`` / ``torch.linalg`` (131072-point FFT
  correlations, dense Cholesky of the Gram matrix, FFT-convolution projections: the route mir_eval itself takes).  It is a
  LIBRARY-built cross-check used by the tests (it also runs on CPU tensors) and is not called by the package.

Neither imports ``oracle/``; tests compare both with the CPU restatement of mir_eval there.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

FLEN = 512  # mir_eval's default distortion-filter length


def _next_pow2(n: int) -> int:
    return 1 << int(np.ceil(np.log2(float(n))))


def _solve(G: torch.Tensor, D: torch.Tensor) -> torch.Tensor:
    """mir_eval: np.linalg.solve, falling back to lstsq when G is singular (e.g. a silent reference).
    G is a Gram matrix (symmetric positive semi-definite), so the batch is first tried with a Cholesky factorisation
    and two triangular solves -- 5x faster than the pivoted LU of torch.linalg.solve for 1024^2 fp64 systems on a
    B200 (tools/time_bss_parts.py) and equal to it to ~1e-12 relative."""
    L, info = torch.linalg.cholesky_ex(G)
    if not bool(info.any()):
        y = torch.linalg.solve_triangular(L, D, upper=False)
        C = torch.linalg.solve_triangular(L.mT, y, upper=True)
        if torch.isfinite(C).all():
            return C
    try:
        C = torch.linalg.solve(G, D)
        if torch.isfinite(C).all():
            return C
    except RuntimeError:
        pass
    return torch.linalg.lstsq(G, D).solution


def _db(num: torch.Tensor, den: torch.Tensor) -> torch.Tensor:
    return torch.where(den == 0, torch.full_like(num, float("inf")), 10.0 * torch.log10(num / den))


def bss_eval_batch(reference: torch.Tensor, estimate: torch.Tensor) -> Dict[str, torch.Tensor]:
    """reference, estimate: [B, N, L] CUDA tensors (float32 or float64; any strides with a contiguous last dimension) ->
    dict of float64 [B, N] tensors 'sdr', 'sir', 'sar' (mir_eval.separation.bss_eval_sources without permutation search,
    512-tap distortion filters) and 'si_sdr' (pb_bss_eval).  Hand-written kernels (csrc/bss_eval.cu), no library call."""
    from . import _lib
    from ._lib import check, lib
    if reference.dim() != 3 or reference.shape != estimate.shape:
        raise ValueError(f"bss_eval_batch: expected matching [B, N, L] tensors, got {tuple(reference.shape)} and {tuple(estimate.shape)}")
    if not (reference.is_cuda and estimate.is_cuda):
        raise _lib.AvsepError("bss_eval_batch: expected CUDA tensors (avsep_b200 has no CPU path); bss_eval_batch_torch is the "
                              "library-built cross-check that also accepts CPU tensors")
    dt = torch.float64 if (reference.dtype == torch.float64 or estimate.dtype == torch.float64) else torch.float32
    ref = reference if reference.dtype == dt else reference.to(dt)
    est = estimate if estimate.dtype == dt else estimate.to(dt)
    if ref.stride(-1) != 1:
        ref = ref.contiguous()
    if est.stride(-1) != 1:
        est = est.contiguous()
    B, N, L = ref.shape
    out = torch.empty((4, B, N), device=ref.device, dtype=torch.float64)
    if B:
        ws = torch.empty(int(lib.avsep_bss_eval_workspace_bytes(B, N, L)), dtype=torch.uint8, device=ref.device)
        check(lib.avsep_bss_eval(ref.data_ptr(), ref.stride(0), ref.stride(1), est.data_ptr(), est.stride(0), est.stride(1),
                                 int(dt == torch.float64), B, N, L, out.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "avsep_bss_eval")
    return dict(sdr=out[0], sir=out[1], sar=out[2], si_sdr=out[3])


def bss_eval_batch_torch(reference: torch.Tensor, estimate: torch.Tensor, flen: int = FLEN) -> Dict[str, torch.Tensor]:
    """LIBRARY-built cross-check of bss_eval_batch (torch.fft / torch.linalg, mir_eval's own route); any device.
    reference, estimate: [B, N, L] -> dict of float64 [B, N] tensors 'sdr', 'sir', 'sar', 'si_sdr'."""
    ref = reference.to(torch.float64)
    est = estimate.to(torch.float64)
    if ref.dim() != 3 or ref.shape != est.shape:
        raise ValueError(f"bss_eval_batch_torch: expected matching [B, N, L] tensors, got {tuple(ref.shape)} and {tuple(est.shape)}")
    B, N, L = ref.shape
    dev = ref.device
    P = L + flen - 1                                   # length of the padded signals / projections
    n_fft = _next_pow2(P)
    sf = torch.fft.rfft(ref, n=n_fft, dim=-1)          # [B, N, F]   (zero padding is implicit in n=)
    sef = torch.fft.rfft(est, n=n_fft, dim=-1)

    k = torch.arange(flen, device=dev)
    toe = (k[None, :] - k[:, None]) % n_fft            # ss[a, c] = ssf[(c - a) mod n_fft]  (scipy.linalg.toeplitz(c, r))
    neg = (-k) % n_fft                                 # D[k]     = ssef[(-k) mod n_fft]

    # reference cross-correlations -> Gram blocks  G[b, i*flen + a, j*flen + c]
    ssf = torch.fft.irfft(sf[:, :, None, :] * sf[:, None, :, :].conj(), n=n_fft, dim=-1)          # [B, N, N, n_fft]
    blocks = ssf[..., toe]                                                                       # [B, N, N, flen, flen]
    G = blocks.permute(0, 1, 3, 2, 4).reshape(B, N * flen, N * flen)
    del ssf
    # reference x estimate cross-correlations -> right-hand sides  D[b, i*flen + k, j]
    ssef = torch.fft.irfft(sf[:, :, None, :] * sef[:, None, :, :].conj(), n=n_fft, dim=-1)        # [B, Nref, Nest, n_fft]
    Dfull = ssef[..., neg]                                                                       # [B, Nref, Nest, flen]
    del ssef
    D = Dfull.permute(0, 1, 3, 2).reshape(B, N * flen, N)

    # projection on all references (e_interf) and on the own reference only (e_spat)
    C = _solve(G, D).reshape(B, N, flen, N)                                                      # [B, ref i, tap, est j]
    Cf = torch.fft.rfft(C.permute(0, 3, 1, 2), n=n_fft, dim=-1)                                  # [B, j, i, F]
    proj_all = torch.fft.irfft((Cf * sf[:, None, :, :]).sum(dim=2), n=n_fft, dim=-1)[..., :P]    # [B, j, P]
    idx = torch.arange(N, device=dev)
    Gd = blocks[:, idx, idx].reshape(B * N, flen, flen)                                          # diagonal blocks (j, j)
    Dd = Dfull[:, idx, idx].reshape(B * N, flen, 1)
    Cd = _solve(Gd, Dd).reshape(B, N, flen)
    proj_own = torch.fft.irfft(torch.fft.rfft(Cd, n=n_fft, dim=-1) * sf, n=n_fft, dim=-1)[..., :P]

    s_true = torch.nn.functional.pad(ref, (0, flen - 1))
    e_spat = proj_own - s_true
    e_interf = proj_all - s_true - e_spat
    e_artif = torch.nn.functional.pad(est, (0, flen - 1)) - s_true - e_spat - e_interf
    s_filt = s_true + e_spat
    en = lambda x: (x * x).sum(dim=-1)
    out = dict(sdr=_db(en(s_filt), en(e_interf + e_artif)),
               sir=_db(en(s_filt), en(e_interf)),
               sar=_db(en(s_filt + e_interf), en(e_artif)))
    # pb_bss_eval.evaluation.si_sdr
    alpha = (ref * est).sum(dim=-1, keepdim=True) / en(ref)[..., None]
    target = alpha * ref
    out["si_sdr"] = 10.0 * torch.log10(en(target) / en(est - target))
    return out


def get_metrics(mix, clean, estimate, sample_rate: int = 11025, metrics_list: Sequence[str] = ("sdr", "sir", "sar", "si_sdr"),
                device: Optional[torch.device] = None, **_ignored) -> Dict[str, float]:
    """Drop-in for the call the reference makes (main.py:260-261): one clip, ``clean`` / ``estimate`` [N, L]
    (numpy or torch), source-averaged floats back.  ``mix`` and ``sample_rate`` are accepted for signature
    compatibility; the input_* metrics asteroid also returns are not used by the reference and not computed."""
    if device is None:
        device = torch.device("cuda")
    c = torch.as_tensor(np.asarray(clean) if not torch.is_tensor(clean) else clean, dtype=torch.float64, device=device)
    e = torch.as_tensor(np.asarray(estimate) if not torch.is_tensor(estimate) else estimate, dtype=torch.float64, device=device)
    res = bss_eval_batch(c[None], e[None])
    return {k: float(res[k].mean()) for k in metrics_list}
