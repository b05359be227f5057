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
