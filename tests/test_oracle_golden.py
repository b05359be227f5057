"""The oracle against the golden vectors produced by the REAL reference code (oracle/make_golden.py).
CPU only; inputs are regenerated from the same seeds the generator used."""
import numpy as np

from conftest import relerr
from oracle import np_oracle as O

TOL = 2e-6   # float32 round-off of the reference's own arithmetic


def _rng_inputs(seed, *shapes):
    g = np.random.default_rng(seed)
    return [g.standard_normal(s).astype(np.float32) for s in shapes]


def test_warpgrid_matches_reference(golden):
    g = golden("warpgrid")
    gw = O.warpgrid(1, 256, 256, True)
    gu = O.warpgrid(1, 512, 256, False)
    assert np.array_equal(gw[0, :, 0, 1], g["warp_y"]) and np.array_equal(gw[0, 0, :, 0], g["warp_x"])
    assert np.array_equal(gu[0, :, 0, 1], g["unwarp_y"]) and np.array_equal(gu[0, 0, :, 0], g["unwarp_x"])
    assert np.array_equal(O.warp_axis_coords(256, True), g["warp_y"])
    assert np.array_equal(O.warp_axis_coords(512, False), g["unwarp_y"])


def test_grid_sample_matches_torch(golden):
    g = golden("grid_sample")
    (xw,) = _rng_inputs(11, (2, 1, 512, 256))
    xw = np.abs(xw)
    assert relerr(O.log_freq_warp(xw, 256, True)[1], g["warp_out_b1"]) < TOL
    (xu,) = _rng_inputs(12, (2, 1, 256, 256))
    assert relerr(O.log_freq_warp(xu, 512, False)[1], g["unwarp_out_b1"]) < TOL
    (xs,) = _rng_inputs(13, (1, 1, 512, 24))
    assert relerr(O.log_freq_warp(xs, 256, True), g["warp_T24"]) < TOL


def test_separable_taps_equal_grid_sample():
    (x,) = _rng_inputs(5, (1, 1, 512, 40))
    ref = O.log_freq_warp(x, 256, True)[0, 0]
    y0, fy = O.axis_taps(256, 512, O.warp_axis_coords(256, True))
    x0, fx = O.axis_taps(40, 40, np.linspace(-1, 1, 40).astype(np.float32))
    xin = np.pad(x[0, 0].astype(np.float64), ((1, 1), (1, 1)))
    rows = (1 - fy)[:, None] * xin[y0 + 1] + fy[:, None] * xin[y0 + 2]
    out = (1 - fx)[None, :] * rows[:, x0 + 1] + fx[None, :] * rows[:, x0 + 2]
    assert relerr(out, ref) < TOL


def test_prepare_matches_reference(golden):
    g = golden("prepare")
    for tag, (binary, weighted) in {"bin_w": (True, True), "ratio_now": (False, False)}.items():
        r = np.random.default_rng(21)
        mags = [np.abs(r.standard_normal((2, 1, 512, 64))).astype(np.float32) * 3 for _ in range(2)]
        mag_mix = (mags[0] + mags[1] * np.float32(0.7)).astype(np.float32)
        o = O.prepare(mag_mix, mags, log_freq=True, weighted_loss=weighted, binary_mask=binary)
        assert relerr(o["mag_mix"], g[f"{tag}_mag_mix"]) < TOL
        assert relerr(o["mags"][0], g[f"{tag}_mags0"]) < TOL
        assert relerr(o["mags"][1], g[f"{tag}_mags1"]) < TOL
        assert relerr(o["log_mag_mix"], g[f"{tag}_log"]) < TOL
        assert relerr(o["weights"], g[f"{tag}_w"]) < TOL
        if binary:
            assert np.mean(o["gt_masks"][0] != g[f"{tag}_gt0"]) < 1e-4
            assert np.mean(o["gt_masks"][1] != g[f"{tag}_gt1"]) < 1e-4
        else:
            assert relerr(o["gt_masks"][0], g[f"{tag}_gt0"]) < TOL


def _synth_inputs():
    B, K, H, W, N = 2, 32, 16, 24, 2
    g = np.random.default_rng(31)
    feat = g.standard_normal((B, K, H, W)).astype(np.float32)
    vecs = (1 / (1 + np.exp(-g.standard_normal((B, N, K))))).astype(np.float32)
    scale = (1 + 0.1 * g.standard_normal(K)).astype(np.float32)
    bias = np.array([0.05], dtype=np.float32)
    gts = [(g.random((B, 1, H, W)) > 0.5).astype(np.float32) for _ in range(N)]
    wgt = np.clip(np.abs(g.standard_normal((B, 1, H, W))), 1e-3, 10).astype(np.float32)
    fi = g.standard_normal((B, K, 3, 5)).astype(np.float32)
    return feat, vecs, scale, bias, gts, wgt, fi


def test_synth_and_criterion_match_reference(golden):
    g = golden("synth")
    feat, vecs, scale, bias, gts, wgt, fi = _synth_inputs()
    assert relerr(O.innerprod_forward(vecs[:, 0], feat, scale, bias), g["z"]) < TOL
    assert relerr(O.innerprod_forward_nosum(vecs[:, 0], feat, scale, bias), g["z_nosum"]) < TOL
    assert relerr(O.innerprod_forward_pixelwise(fi, feat, scale, bias), g["z_pixelwise"]) < TOL
    assert relerr(O.innerprod_forward(vecs[:, 1], feat, None, bias), g["z_bias"]) < TOL
    for kind in ("bce", "l1", "l2"):
        o = O.synth_loss_and_grads(kind, feat, vecs, scale, bias, gts, wgt)
        assert abs(o["loss"] - float(g[f"{kind}_loss"])) / abs(float(g[f"{kind}_loss"])) < TOL
        assert relerr(o["dfeat"], g[f"{kind}_dfeat"]) < TOL
        assert relerr(o["dvecs"], g[f"{kind}_dvecs"]) < TOL
        assert relerr(o["dscale"], g[f"{kind}_dscale"]) < TOL
        assert abs(o["dbias"] - float(g[f"{kind}_dbias"][0])) / abs(float(g[f"{kind}_dbias"][0])) < TOL
    assert relerr(O.synth_loss_and_grads("bce", feat, vecs, scale, bias, gts, wgt)["preds"][1], g["pred1"]) < TOL


def test_bce_saturation_semantics(golden):
    """z = +-40: the loss term is 100 * w and the gradient is exactly zero."""
    g = golden("synth")
    z, t = g["sat_z"], g["sat_t"]
    p = O.sigmoid_f32(z)
    w = np.full(5, 2.0)
    loss = float(np.sum(w * O.loss_elem("bce", p, t)))
    assert abs(loss - float(g["sat_loss"])) / float(g["sat_loss"]) < TOL
    dp = O.loss_grad_wrt_pred("bce", p, t, w, 1, 1)
    grad = dp * p.astype(np.float64) * (1 - p.astype(np.float64))
    assert np.allclose(grad, g["sat_grad"], rtol=1e-5, atol=1e-12)
    assert grad[0] == 0.0 and g["sat_grad"][0] == 0.0


def test_pit_matches_reference(golden):
    g = golden("pit")
    for S in (2, 3):
        r = np.random.default_rng(40 + S)
        pr = r.random((3, 12, 10, S)).astype(np.float32) * 0.98 + 0.01
        tg = (r.random((3, 12, 10, S)) > 0.5).astype(np.float32)
        pr[1] = np.clip(tg[1][..., ::-1] * 0.9 + 0.05, 0.01, 0.99)
        ww = np.repeat(np.abs(r.standard_normal((3, 12, 10, 1))).astype(np.float32), S, axis=-1)
        loss, perms = O.pit("bce", pr, tg, ww)
        assert relerr(loss, g[f"S{S}_loss"]) < TOL
        assert np.array_equal(np.asarray(perms), g[f"S{S}_perms"])
        assert tuple(perms[1]) != tuple(range(S))
        assert relerr(O.pit_grad("bce", pr, tg, ww, perms, np.array([1.0, 2.0, 0.5])), g[f"S{S}_grad"]) < TOL


def test_pit_doctest_property():
    """criterion.py:96-104: a permuted copy of the targets has zero PIT loss under MSE."""
    r = np.random.default_rng(0)
    targets = r.random((2, 32, 4))
    p = (3, 0, 2, 1)
    preds = targets[..., p]
    loss, perms = O.pit("l2", preds, targets, np.ones_like(targets))
    assert np.allclose(loss, 0.0)


def test_stft_matches_torch_stft(golden):
    g = golden("stft")
    src = O.synth_sources(1, 2, seed=1234)[0]
    mix = (src[0] + src[1]).astype(np.float32)
    amp, ph = O.stft_magphase(mix)
    fr = g["frames"]
    assert amp.shape == (512, 256) and O.n_frames(65535) == 256
    assert relerr(amp[:, fr], g["amp_mix"]) < TOL
    spec_o = amp[:, fr].astype(np.float64) * np.exp(1j * ph[:, fr].astype(np.float64))
    spec_g = g["amp_mix"].astype(np.float64) * np.exp(1j * g["phase_mix"].astype(np.float64))
    assert np.max(np.abs(spec_o - spec_g)) / np.max(np.abs(spec_g)) < TOL
    assert relerr(O.stft_magphase(src[0])[0][:, fr], g["amp_src0"]) < TOL
    short = src[1][:3000].copy()
    a_s, _ = O.stft_magphase(short)
    assert a_s.shape == g["amp_short"].shape and relerr(a_s, g["amp_short"]) < TOL


def test_istft_matches_torch_istft(golden):
    g = golden("stft")
    src = O.synth_sources(1, 2, seed=1234)[0]
    mix = (src[0] + src[1]).astype(np.float32)
    amp, ph = O.stft_magphase(mix)
    mask = (np.random.default_rng(51).random((512, 256)) > 0.4).astype(np.float32)
    wav = O.istft_reconstruction(amp * mask, ph)
    assert wav.shape[0] == int(g["istft_len"]) == 65280
    assert relerr(wav[:4096], g["istft_masked_head"]) < 5e-6
    assert relerr(wav[30000:34096], g["istft_masked_mid"]) < 5e-6
    assert relerr(wav[-4096:], g["istft_masked_tail"]) < 5e-6
    rt = O.istft_reconstruction(amp, ph)
    assert np.max(np.abs(rt - mix[:65280])) < 2e-6        # round trip (SURVEY.md 8(a) row 9 probe)


def test_stft_and_istft_match_scipy_signal():
    """Second independent implementation beside torch.stft / torch.istft: scipy.signal.stft / istft with librosa's conventions
    (librosa itself takes its window from scipy.signal.get_window): periodic Hann of 1022, hop 256, reflect ('even') centre
    padding, no scaling.  librosa is not installable here, so the oracle's STFT / iSTFT restatement (dataset/base.py:142-147,
    utils.py:101-104) is pinned by two unrelated code bases instead."""
    import scipy.signal
    rng = np.random.default_rng(11)
    win_sum = O.hann_periodic(1022).sum()
    for length in (65535, 40000, 3000):
        x = (rng.standard_normal(length) * 0.3).astype(np.float32)
        spec = O.stft_complex(x)
        _, _, z = scipy.signal.stft(x.astype(np.float64), window="hann", nperseg=1022, noverlap=1022 - 256, nfft=1022, detrend=False,
                                    return_onesided=True, boundary="even", padded=False, scaling="spectrum")
        z = z * win_sum                                              # scipy divides by sum(window), librosa does not scale
        assert z.shape == spec.shape == (512, 1 + length // 256)
        assert np.max(np.abs(z - spec)) / np.max(np.abs(spec)) < 1e-12
        # inverse: librosa.istft(spec, hop_length=256) -> hop * (T - 1) samples, window-sum-square normalised, centre trimmed
        rec = O.istft(spec)
        _, y = scipy.signal.istft(spec / win_sum, window="hann", nperseg=1022, noverlap=1022 - 256, nfft=1022, input_onesided=True,
                                  boundary=True, scaling="spectrum")
        n = min(rec.shape[0], y.shape[0])
        assert rec.shape[0] == 256 * (spec.shape[1] - 1) and abs(rec.shape[0] - y.shape[0]) <= 1022
        assert np.max(np.abs(rec[:n] - y[:n])) < 1e-10
        assert np.max(np.abs(rec[:n] - x[:n])) < 1e-6               # both reconstruct the input


def test_calc_metrics_flow_matches_reference(golden):
    from oracle import bss_eval as BSS
    g = golden("calc_metrics")
    waves = O.synth_sources(2, 2, seed=77)
    mixw = waves.sum(axis=1).astype(np.float32)
    amp_m, ph_m = zip(*[O.stft_magphase(mixw[j]) for j in range(2)])
    amp_m, ph_m = np.stack(amp_m)[:, None], np.stack(ph_m)[:, None]
    pm = [g["pred_mask0"].astype(np.float32), g["pred_mask1"].astype(np.float32)]
    mix_o, pred_o, _ = O.separate(amp_m, ph_m, pm)
    j = 1
    m = BSS.get_metrics(mix_o[j], waves[j][:, :mix_o.shape[1]], pred_o[j])
    # masks were stored as float16, so thresholded masks can flip on a handful of bins: 0.01 dB budget
    assert abs(m["sdr"] - g["sdr"][j]) < 0.01 and abs(m["sir"] - g["sir"][j]) < 0.01
    assert abs(m["si_sdr"] - g["si_sdr"][j]) < 0.01
