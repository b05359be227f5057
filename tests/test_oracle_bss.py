"""Closed-form properties of the BSS-eval restatement (mir_eval / pb_bss_eval are absent: parity unpinned)."""
import numpy as np

from oracle import bss_eval as BSS


def _srcs(n=2, L=4000, seed=0):
    g = np.random.default_rng(seed)
    t = np.arange(L) / 11025.0
    return np.stack([np.sin(2 * np.pi * (220 * (k + 1)) * t) * (0.5 + 0.1 * k) + 0.01 * g.standard_normal(L)
                     for k in range(n)])


def test_perfect_estimate_has_huge_sdr():
    s = _srcs()
    sdr, sir, sar = BSS.bss_eval_sources(s, s.copy())
    assert np.all(sdr > 100) and np.all(sir > 100) and np.all(sar > 100)


def test_filtered_source_is_not_distortion():
    """BSS-eval v3 allows a 512-tap filtering of the target: a delayed, scaled copy scores as clean."""
    s = _srcs()
    s[:, -64:] = 0.0                      # so the 7-sample delay does not push energy past the end
    est = np.stack([0.5 * np.concatenate([np.zeros(7), s[0][:-7]]), s[1]])
    sdr, _, _ = BSS.bss_eval_sources(s, est)
    assert sdr[0] > 60


def test_additive_noise_sdr():
    s = _srcs(seed=3)
    g = np.random.default_rng(9)
    noise = 0.05 * g.standard_normal(s.shape)
    sdr, sir, sar = BSS.bss_eval_sources(s, s + noise)
    expect = 10 * np.log10(np.sum(s ** 2, axis=1) / np.sum(noise ** 2, axis=1))
    assert np.all(np.abs(sdr - expect) < 1.5)      # projection removes a little of the noise
    assert np.all(sir > sdr)


def test_interference_shows_in_sir():
    s = _srcs(seed=4)
    est = np.stack([s[0] + 0.3 * s[1], s[1]])
    sdr, sir, sar = BSS.bss_eval_sources(s, est)
    assert abs(sir[0] - 10 * np.log10(np.sum(s[0] ** 2) / np.sum((0.3 * s[1]) ** 2))) < 0.5
    assert sar[0] > 40


def test_si_sdr_scale_invariant():
    s = _srcs(seed=5)
    g = np.random.default_rng(1)
    est = s + 0.1 * g.standard_normal(s.shape)
    a = BSS.si_sdr(s, est)
    b = BSS.si_sdr(s, 3.7 * est)
    assert np.allclose(a, b, atol=1e-9)
    m = BSS.get_metrics(s.sum(0), s, est)
    assert set(m) == {"sdr", "sir", "sar", "si_sdr"} and abs(m["si_sdr"] - a.mean()) < 1e-12


def _project_time_domain(refs, est, flen):
    """Independent restatement of the BSS-eval projection straight from its definition (Vincent et al. 2006): least
    squares onto the explicitly built matrix of delayed references, no FFT, no Toeplitz Gram matrix."""
    nsrc, L = refs.shape
    P = L + flen - 1
    A = np.zeros((P, nsrc * flen))
    for i in range(nsrc):
        for tau in range(flen):
            A[tau:tau + L, i * flen + tau] = refs[i]
    y = np.concatenate([est, np.zeros(flen - 1)])
    c, *_ = np.linalg.lstsq(A, y, rcond=None)
    return A @ c


def _bss_time_domain(refs, ests, flen):
    nsrc, L = refs.shape
    out = []
    for j in range(nsrc):
        s_true = np.concatenate([refs[j], np.zeros(flen - 1)])
        e_spat = _project_time_domain(refs[j:j + 1], ests[j], flen) - s_true
        e_interf = _project_time_domain(refs, ests[j], flen) - s_true - e_spat
        e_artif = np.concatenate([ests[j], np.zeros(flen - 1)]) - s_true - e_spat - e_interf
        s_filt = s_true + e_spat
        db = lambda a, b: 10 * np.log10(np.sum(a ** 2) / np.sum(b ** 2))
        out.append((db(s_filt, e_interf + e_artif), db(s_filt, e_interf), db(s_filt + e_interf, e_artif)))
    return np.array(out).T


def test_fft_toeplitz_route_equals_time_domain_least_squares():
    """The restatement (FFT cross-correlations -> block-Toeplitz Gram system -> FFT-convolution projections, the route
    mir_eval takes) against an independent time-domain least-squares projection on short clips: <= 1e-6 dB."""
    for nsrc, L, flen, seed in ((2, 700, 24, 0), (3, 900, 16, 1), (2, 1500, 64, 2)):
        g = np.random.default_rng(seed)
        t = np.arange(L) / 11025.0
        refs = np.stack([np.sin(2 * np.pi * (300 + 170 * k) * t + k) * (0.4 + 0.1 * k) + 0.05 * g.standard_normal(L) for k in range(nsrc)])
        mixm = np.eye(nsrc) + 0.25 * g.standard_normal((nsrc, nsrc))
        ests = mixm @ refs + 0.08 * g.standard_normal((nsrc, L))
        ests[0, 5:] += 0.3 * refs[0, :-5]                       # a delayed copy: allowed distortion, exercises the filter taps
        a = np.array(BSS.bss_eval_sources(refs, ests, flen=flen))
        b = _bss_time_domain(refs, ests, flen)
        assert np.max(np.abs(a - b)) < 1e-6, (nsrc, L, flen, np.max(np.abs(a - b)))
