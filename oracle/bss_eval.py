"""CPU oracle for the separation metrics the reference obtains from asteroid.

TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py header).

Reference call site: main.py:260-261
    get_metrics(mix_wav[0:L], gts_wav[N,L], preds_wav[N,L], sample_rate=11025,
                metrics_list=['sdr','sir','sar','si_sdr'])
``asteroid`` (version unpinned by the reference, not installable here) routes
'sdr'/'sir'/'sar' to ``mir_eval.separation.bss_eval_sources(reference, estimate,
compute_permutation=False)`` through ``pb_bss_eval`` and 'si_sdr' to
``pb_bss_eval.evaluation.si_sdr``; results are averaged over sources
(``average=True`` default).  Both algorithms are restated below from their
published definitions (Vincent et al. 2006 BSS-eval v3 with 512-tap distortion
filters; Le Roux et al. 2019 SI-SDR).  Parity status: **unpinned** -- neither
library is present, so these functions are checked only against closed-form
properties (tests/test_oracle_bss.py), not against library output.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.linalg
import scipy.signal

FLEN = 512  # mir_eval default distortion-filter length


def _project(reference_sources: np.ndarray, estimated_source: np.ndarray, flen: int = FLEN) -> np.ndarray:
    """Least-squares projection of the estimate onto delayed versions of the references
    (mir_eval.separation._project)."""
    nsrc, nsampl = reference_sources.shape
    ref = np.hstack((reference_sources, np.zeros((nsrc, flen - 1))))
    est = np.hstack((estimated_source, np.zeros(flen - 1)))
    n_fft = int(2 ** np.ceil(np.log2(nsampl + flen - 1.0)))
    sf = scipy.fft.fft(ref, n=n_fft, axis=1)
    sef = scipy.fft.fft(est, n=n_fft)
    G = np.zeros((nsrc * flen, nsrc * flen))
    for i in range(nsrc):
        for j in range(nsrc):
            ssf = np.real(scipy.fft.ifft(sf[i] * np.conj(sf[j])))
            ss = scipy.linalg.toeplitz(np.hstack((ssf[0], ssf[-1:-flen:-1])), r=ssf[:flen])
            G[i * flen:(i + 1) * flen, j * flen:(j + 1) * flen] = ss
            G[j * flen:(j + 1) * flen, i * flen:(i + 1) * flen] = ss.T
    D = np.zeros(nsrc * flen)
    for i in range(nsrc):
        ssef = np.real(scipy.fft.ifft(sf[i] * np.conj(sef)))
        D[i * flen:(i + 1) * flen] = np.hstack((ssef[0], ssef[-1:-flen:-1]))
    try:
        C = np.linalg.solve(G, D).reshape(flen, nsrc, order="F")
    except np.linalg.LinAlgError:
        C = np.linalg.lstsq(G, D, rcond=None)[0].reshape(flen, nsrc, order="F")
    sproj = np.zeros(nsampl + flen - 1)
    for i in range(nsrc):
        sproj += scipy.signal.fftconvolve(C[:, i], ref[i])[:nsampl + flen - 1]
    return sproj


def _safe_db(num: float, den: float) -> float:
    if den == 0:
        return np.inf
    return 10 * np.log10(num / den)


def bss_eval_sources(reference_sources: np.ndarray, estimated_sources: np.ndarray, flen: int = FLEN):
    """mir_eval.separation.bss_eval_sources(..., compute_permutation=False) -> (sdr, sir, sar) [nsrc]."""
    reference_sources = np.atleast_2d(np.asarray(reference_sources, dtype=np.float64))
    estimated_sources = np.atleast_2d(np.asarray(estimated_sources, dtype=np.float64))
    nsrc, nsampl = estimated_sources.shape
    sdr = np.empty(nsrc)
    sir = np.empty(nsrc)
    sar = np.empty(nsrc)
    for j in range(nsrc):
        s_true = np.hstack((reference_sources[j], np.zeros(flen - 1)))
        e_spat = _project(reference_sources[j, np.newaxis, :], estimated_sources[j], flen) - s_true
        e_interf = _project(reference_sources, estimated_sources[j], flen) - s_true - e_spat
        e_artif = -s_true - e_spat - e_interf
        e_artif[:nsampl] += estimated_sources[j]
        s_filt = s_true + e_spat
        sdr[j] = _safe_db(np.sum(s_filt ** 2), np.sum((e_interf + e_artif) ** 2))
        sir[j] = _safe_db(np.sum(s_filt ** 2), np.sum(e_interf ** 2))
        sar[j] = _safe_db(np.sum((s_filt + e_interf) ** 2), np.sum(e_artif ** 2))
    return sdr, sir, sar


def si_sdr(reference: np.ndarray, estimation: np.ndarray) -> np.ndarray:
    """pb_bss_eval.evaluation.si_sdr: 10 log10(|a s|^2 / |s_hat - a s|^2), a = <s, s_hat>/|s|^2."""
    estimation, reference = np.broadcast_arrays(np.asarray(estimation, dtype=np.float64),
                                                np.asarray(reference, dtype=np.float64))
    ref_energy = np.sum(reference ** 2, axis=-1, keepdims=True)
    alpha = np.sum(reference * estimation, axis=-1, keepdims=True) / ref_energy
    proj = alpha * reference
    noise = estimation - proj
    return 10 * np.log10(np.sum(proj ** 2, axis=-1) / np.sum(noise ** 2, axis=-1))


def get_metrics(mix: np.ndarray, clean: np.ndarray, estimate: np.ndarray) -> dict:
    """asteroid.metrics.get_metrics(..., metrics_list=['sdr','sir','sar','si_sdr'], average=True)
    as used at main.py:260-266: source-averaged output metrics (input_* keys omitted: unused)."""
    sdr, sir, sar = bss_eval_sources(clean, estimate)
    return dict(sdr=float(np.mean(sdr)), sir=float(np.mean(sir)), sar=float(np.mean(sar)),
                si_sdr=float(np.mean(si_sdr(clean, estimate))))
