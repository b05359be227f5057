"""GPU parity of the fused end-to-end steps (what bench.py times) and of the calc_metrics drop-in,
including the SDR / SIR / SAR criterion of the north star (per clip within 0.01 dB of the oracle)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import relerr
from oracle import bss_eval as BSS
from oracle import np_oracle as O

import avsep_b200 as av

pytestmark = pytest.mark.gpu
TOL = 1e-4
L = 65535


def _inputs(B, N, K=32, seed=1234):
    src = O.synth_sources(B, N, seed=seed)                      # [B, N, L]
    mix = src.sum(axis=1).astype(np.float32)
    waves = np.concatenate([mix[None], np.transpose(src, (1, 0, 2))], axis=0)   # [N+1, B, L]
    g = np.random.default_rng(seed + 1)
    feat = g.standard_normal((B, K, 256, 256)).astype(np.float32)
    vecs = (1 / (1 + np.exp(-g.standard_normal((B, N, K))))).astype(np.float32)
    scale = (1 + 0.05 * g.standard_normal(K)).astype(np.float32)
    bias = np.array([0.02], dtype=np.float32)
    return np.ascontiguousarray(waves), feat, vecs, scale, bias


def _oracle_front(waves, N, B):
    amp = np.stack([np.stack([O.stft_magphase(waves[j, b])[0] for b in range(B)]) for j in range(N + 1)])  # [N+1,B,512,256]
    prep = O.prepare(amp[0][:, None], [amp[1 + n][:, None] for n in range(N)])
    return amp, prep


def test_train_step_vs_oracle():
    B, N, K = 2, 2, 32
    waves, feat, vecs, scale, bias = _inputs(B, N, K)
    amp, prep = _oracle_front(waves, N, B)
    st = av.Step(B, N, K, L)
    c = lambda a: torch.from_numpy(a).cuda()
    dfeat = torch.empty((B, K, 256, 256), device="cuda")
    dvecs = torch.empty((B, N, K), device="cuda")
    dscale = torch.empty(K, device="cuda")
    dbias = torch.empty(1, device="cuda")
    loss = torch.empty(1, device="cuda")
    tfeat = c(feat)
    st.train_dev(c(waves), c(vecs), tfeat, c(scale), c(bias), dfeat, dvecs, dscale, dbias, loss)
    torch.cuda.synchronize()
    # front end
    assert relerr(st.buffer("mag").cpu().numpy().reshape(N + 1, B, 512, 256), amp) < TOL
    assert relerr(st.buffer("mag_mix_w").cpu().numpy(), prep["mag_mix"]) < TOL
    # log magnitude: per-element relative error of |X| is unbounded at spectral nulls, so the
    # log is checked against the GPU's own warped magnitude, and against the oracle where the bin is not a null
    mixw_gpu = st.buffer("mag_mix_w").cpu().numpy()
    log_gpu = st.buffer("log_mag_mix").cpu().numpy()
    assert np.max(np.abs(log_gpu - np.log(mixw_gpu.astype(np.float64)))) < 1e-5
    loud = prep["mag_mix"] > 1e-2 * prep["mag_mix"].max()
    assert np.max(np.abs(log_gpu[loud] - prep["log_mag_mix"][loud])) < 1e-3
    assert relerr(st.buffer("weights").cpu().numpy(), prep["weights"]) < TOL
    gt_gpu = st.buffer("gt_masks").cpu().numpy()
    mism = np.mean(gt_gpu != np.stack(prep["gt_masks"]))
    assert mism < 1e-3, mism                                      # ties of m > 0.5 * mix at the 4e-6 STFT error level
    # back end on the GPU's own gt/weights (so a flipped tie does not leak into the comparison)
    o = O.synth_loss_and_grads("bce", feat, vecs, scale, bias, [gt_gpu[n] for n in range(N)],
                               st.buffer("weights").cpu().numpy())
    assert abs(float(loss) - o["loss"]) / o["loss"] < TOL
    assert relerr(dfeat.cpu().numpy(), o["dfeat"]) < TOL
    assert relerr(dvecs.cpu().numpy(), o["dvecs"]) < TOL
    assert relerr(dscale.cpu().numpy(), o["dscale"]) < TOL
    assert abs(float(dbias) - o["dbias"]) / abs(o["dbias"]) < TOL
    # and end to end against the pure-oracle loss (ties included): still within tolerance
    o2 = O.synth_loss_and_grads("bce", feat, vecs, scale, bias, prep["gt_masks"], prep["weights"])
    assert abs(float(loss) - o2["loss"]) / o2["loss"] < 1e-3
    # host-buffer entry point gives the same numbers
    wh = torch.from_numpy(waves).pin_memory()
    vh = torch.from_numpy(vecs).pin_memory()
    dfeat2 = torch.empty_like(dfeat)
    dv_host = torch.empty((B, N, K)).pin_memory()
    lh = st.train_host(wh, vh, tfeat, c(scale), c(bias), dfeat2, dv_host)
    assert lh == float(loss) and torch.equal(dfeat, dfeat2) and torch.equal(dv_host.cuda(), dvecs)
    # CUDA-graph replay of the step (side stream): same bits as the plain launches
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        st.set_graph(True)
        dfeat_g = torch.empty_like(dfeat)
        loss_g = torch.empty(1, device="cuda")
        tw, tv, tsc, tb = c(waves), c(vecs), c(scale), c(bias)
        before = av.launch_count()
        for _ in range(4):                                         # call 1 direct, call 2 captures + launches, calls 3-4 replay
            dfeat_g.zero_()
            st.train_dev(tw, tv, tfeat, tsc, tb, dfeat_g, None, None, None, loss_g)
        side.synchronize()
        assert av.launch_count() - before == 4 * 4                 # dft_fold3 (staging fused in), prepare, synth_train, finalize per step
        assert float(loss_g) == float(loss) and torch.equal(dfeat_g, dfeat)
        st.set_graph(False)
    torch.cuda.current_stream().wait_stream(side)
    # separate forward / backward kernels (two passes over feat_sound) agree with the fused training kernel
    st.set_fused_train(False)
    loss4 = torch.empty(1, device="cuda")
    dfeat4, dvecs4, dscale4, dbias4 = torch.empty_like(dfeat), torch.empty_like(dvecs), torch.empty_like(dscale), torch.empty_like(dbias)
    st.train_dev(c(waves), c(vecs), tfeat, c(scale), c(bias), dfeat4, dvecs4, dscale4, dbias4, loss4)
    torch.cuda.synchronize()
    assert abs(float(loss4) - float(loss)) < 1e-6 * abs(float(loss))
    assert relerr(dfeat4.cpu().numpy(), dfeat.cpu().numpy()) < 1e-5 and relerr(dvecs4.cpu().numpy(), dvecs.cpu().numpy()) < 1e-5
    assert relerr(dscale4.cpu().numpy(), dscale.cpu().numpy()) < 1e-5 and abs(float(dbias4) - float(dbias)) < 1e-5 * abs(float(dbias))
    st.set_fused_train(True)
    # mixture formed on the GPU from the N sources (dataset/music.py:127): same bits for N = 2 (one IEEE add)
    st.set_mix_on_device(True)
    st.plan.set_mix_linear(False)                                # the summed waveform goes through its own transform
    mag_ref = st.buffer("mag")
    gt_ref = st.buffer("gt_masks")
    loss3 = torch.empty(1, device="cuda")
    dfeat3 = torch.empty_like(dfeat)
    st.train_dev(c(np.ascontiguousarray(waves[1:])), c(vecs), tfeat, c(scale), c(bias), dfeat3, None, None, None, loss3)
    torch.cuda.synchronize()
    assert float(loss3) == float(loss) and torch.equal(dfeat3, dfeat)
    lh2 = st.train_host(torch.from_numpy(np.ascontiguousarray(waves[1:])).pin_memory(), vh, tfeat, c(scale), c(bias), dfeat2, dv_host)
    assert lh2 == float(loss)
    # default: the STFT is linear, so the mixture spectrum is X0 + X1 and only the two sources are transformed
    # (dft_fold3m).  Same magnitudes to fp32 rounding; a binary ground-truth mask may flip only where
    # mag_n / mag_mix sits within that rounding of 0.5.
    st.plan.set_mix_linear(True)
    loss5 = torch.empty(1, device="cuda")
    dfeat5 = torch.empty_like(dfeat)
    st.train_dev(c(np.ascontiguousarray(waves[1:])), c(vecs), tfeat, c(scale), c(bias), dfeat5, None, None, None, loss5)
    torch.cuda.synchronize()
    mag5 = st.buffer("mag")
    assert float((mag5 - mag_ref).abs().max()) / float(mag_ref.max()) < 2e-5
    assert torch.equal(mag5[B:], mag_ref[B:]) or float((mag5[B:] - mag_ref[B:]).abs().max()) / float(mag_ref.max()) < 2e-6
    for b in range(B):
        assert relerr(mag5[b].cpu().numpy(), O.stft_magphase(waves[0, b])[0]) < TOL          # mixture magnitude vs the oracle
    flipped = int((st.buffer("gt_masks") != gt_ref).sum())
    assert flipped <= 2e-4 * gt_ref.numel(), flipped
    assert abs(float(loss5) - float(loss)) < 1e-4 * abs(float(loss))
    same = (st.buffer("gt_masks") == gt_ref).all(dim=0)                                       # [B,1,256,T] pixels with identical targets
    d5, d0 = dfeat5 * same, dfeat * same
    assert float((d5 - d0).abs().max()) / float(d0.abs().max()) < 1e-4
    st.set_mix_on_device(False)
    st.close()


def test_pcm16_handoff_is_bit_identical_to_the_float_handoff():
    """Opt-in 16-bit PCM hand-off (dataset/base.py:156-172, dataset/music.py:120): int16 / 32768 -> * gain -> clip -> / N on
    the device, in the loader's fp32 order, equals numpy's float32 result bit for bit, and the train step fed with PCM
    equals the step fed with that float waveform bit for bit."""
    B, N, K = 2, 2, 32
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32768, size=(N, B, L), dtype=np.int16)
    pcm[0, 0, :100] = 32767                                          # exercise the clip after the gain
    pcm[1, 1, :100] = -32768
    gains = (rng.random((N, B)) + 0.5).astype(np.float32)            # base.py:165-167: random.random() + 0.5
    # the loader's arithmetic in numpy float32
    audio = pcm.astype(np.float32) / np.float32(32768.0)
    audio = audio * gains[:, :, None]
    audio[audio > 1.] = 1.
    audio[audio < -1.] = -1.
    audio = (audio / N).astype(np.float32)
    got = av.pcm16_to_wave(torch.from_numpy(pcm).cuda(), torch.from_numpy(gains).cuda(), num_mix=N).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, audio)
    assert np.array_equal(av.pcm16_to_wave(torch.from_numpy(pcm).cuda(), None, num_mix=3).cpu().numpy(),
                          (np.clip(pcm.astype(np.float32) / np.float32(32768.0), -1, 1) / 3).astype(np.float32))
    # the whole train step: PCM in vs the float waveform in
    _, feat, vecs, scale, bias = _inputs(B, N, K, seed=77)
    c = lambda a: torch.from_numpy(a).cuda()
    tfeat, tscale, tbias = c(feat), c(scale), c(bias)
    vecs_host = torch.from_numpy(vecs).pin_memory()
    res = {}
    for mode in ("pcm16", "float"):
        st = av.Step(B, N, K, L)
        st.set_mix_on_device(True)
        dfeat = torch.empty((B, K, 256, 256), device="cuda")
        loss_host = torch.zeros(1).pin_memory()
        dvecs_host = torch.empty((B, N, K)).pin_memory()
        if mode == "pcm16":
            st.train_pcm16_async(torch.from_numpy(pcm).pin_memory(), torch.from_numpy(gains).pin_memory(), vecs_host, tfeat, tscale, tbias,
                                 dfeat, loss_host, dvecs_host)
        else:
            st.train_host_async(torch.from_numpy(audio).pin_memory(), vecs_host, tfeat, tscale, tbias, dfeat, loss_host, dvecs_host)
        torch.cuda.synchronize()
        res[mode] = (float(loss_host[0]), st.buffer("mag").clone(), dfeat.clone(), dvecs_host.clone())
        st.close()
    assert res["pcm16"][0] == res["float"][0] and res["float"][0] > 0
    assert torch.equal(res["pcm16"][1], res["float"][1]) and torch.equal(res["pcm16"][2], res["float"][2])
    assert torch.equal(res["pcm16"][3], res["float"][3])
    st = av.Step(B, N, K, L)
    with pytest.raises(av.AvsepError):                               # float buffer handed to the PCM entry point
        st.train_pcm16_async(torch.zeros((N + 1, B, L)).pin_memory(), None, vecs_host, tfeat, tscale, tbias, None, torch.zeros(1).pin_memory())
    st.close()


def test_eval_host_async_equals_eval_dev():
    """avsep_step_eval_host_async: host waveforms in, separated waveforms out (or left on the device), no sync inside."""
    B, N, K = 2, 2, 32
    waves, feat, vecs, scale, bias = _inputs(B, N, K, seed=5)
    c = lambda a: torch.from_numpy(a).cuda()
    st = av.Step(B, N, K, L)
    wav_dev = torch.empty((B, N + 1, 65280), device="cuda")
    st.eval_dev(c(waves), c(vecs), c(feat), c(scale), c(bias), wav_dev)
    wav_host = torch.empty((B, N + 1, 65280)).pin_memory()
    st.eval_host_async(torch.from_numpy(waves).pin_memory(), torch.from_numpy(vecs).pin_memory(), c(feat), c(scale), c(bias), wav_host)
    torch.cuda.synchronize()
    assert torch.equal(wav_host, wav_dev.cpu())
    st.eval_host_async(torch.from_numpy(waves).pin_memory(), torch.from_numpy(vecs).pin_memory(), c(feat), c(scale), c(bias), None)
    torch.cuda.synchronize()
    assert torch.equal(st.view("wav"), wav_dev)
    # CUDA-graph replay of the eval step (side stream): same bits as the plain launches, same launch count per step
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        tw, tv, tf, tsc, tb = c(waves), c(vecs), c(feat), c(scale), c(bias)
        wav_g = torch.empty_like(wav_dev)
        before = av.launch_count()
        st.eval_dev(tw, tv, tf, tsc, tb, wav_g)
        per_step = av.launch_count() - before
        st.set_graph(True)
        before = av.launch_count()
        for _ in range(4):                                         # call 1 direct, call 2 captures + launches, calls 3-4 replay
            wav_g.zero_()
            st.eval_dev(tw, tv, tf, tsc, tb, wav_g)
        side.synchronize()
        assert av.launch_count() - before == 4 * per_step
        assert torch.equal(wav_g, wav_dev)
        st.eval_dev(tw, tv, tf, tsc, tb, wav_g, binary_mask=False)   # a changed argument tuple is not replayed from the old graph
        side.synchronize()
        assert not torch.equal(wav_g, wav_dev)
        st.set_graph(False)
    torch.cuda.current_stream().wait_stream(side)
    st.close()


def test_step_rejects_wrong_waveform_layout():
    """The C-ABI takes raw pointers, so the Python wrapper checks that the waveform tensor matches the step's input
    layout ([N+1, B, L] with the mixture, or [N, B, L] sources after set_mix_on_device) instead of reading past its end."""
    B, N, K = 1, 2, 32
    waves, feat, vecs, scale, bias = _inputs(B, N, K, seed=5)
    c = lambda a: torch.from_numpy(a).cuda()
    st = av.Step(B, N, K, L)
    dfeat = torch.empty((B, K, 256, 256), device="cuda")
    with pytest.raises(av.AvsepError):
        st.train_dev(c(np.ascontiguousarray(waves[1:])), c(vecs), c(feat), c(scale), c(bias), dfeat)      # sources only, mixture expected
    st.set_mix_on_device(True)
    with pytest.raises(av.AvsepError):
        st.train_dev(c(waves), c(vecs), c(feat), c(scale), c(bias), dfeat)                               # mixture given, sources expected
    with pytest.raises(av.AvsepError):
        st.train_host(torch.from_numpy(waves).double(), torch.from_numpy(vecs), c(feat), c(scale), c(bias), dfeat)
    loss = torch.empty(1, device="cuda")
    st.train_dev(c(np.ascontiguousarray(waves[1:])), c(vecs), c(feat), c(scale), c(bias), dfeat, None, None, None, loss)
    torch.cuda.synchronize()
    assert np.isfinite(float(loss))
    st.close()


def test_eval_step_vs_oracle_with_sdr():
    B, N, K = 2, 2, 32
    waves, feat, vecs, scale, bias = _inputs(B, N, K, seed=4321)
    c = lambda a: torch.from_numpy(a).cuda()
    # make the synthesizer emit useful masks: feat channel 0 carries the logit of the ideal mask of source 0
    amp, prep = _oracle_front(waves, N, B)
    feat[:] *= 0.05
    logit = (prep["gt_masks"][0][:, 0] * 2 - 1) * 3.0
    feat[:, 0] = logit
    feat[:, 1] = -logit
    vecs[:] = 0.0
    vecs[:, 0, 0] = 1.0
    vecs[:, 1, 1] = 1.0
    scale[:] = 1.0
    bias[:] = 0.0
    st = av.Step(B, N, K, L)
    wav = torch.empty((B, N + 1, 65280), device="cuda")
    st.eval_dev(c(waves), c(vecs), c(feat), c(scale), c(bias), wav, binary_mask=True, mask_thres=0.5)
    torch.cuda.synchronize()
    pred = st.buffer("pred_masks").cpu().numpy()                 # [N,B,1,256,256]
    z = [O.innerprod_forward(vecs[:, n], feat, scale, bias) for n in range(N)]
    for n in range(N):
        assert relerr(pred[n], O.sigmoid_f32(z[n].astype(np.float32))) < TOL
    phase = np.stack([O.stft_magphase(waves[0, b])[1] for b in range(B)])[:, None]
    mix_o, pred_o, _ = O.separate(amp[0][:, None], phase, [pred[n] for n in range(N)])
    got = wav.cpu().numpy().astype(np.float64)
    scale_w = np.max(np.abs(mix_o))
    assert np.max(np.abs(got[:, 0] - mix_o)) / scale_w < TOL
    for n in range(N):
        assert np.max(np.abs(got[:, 1 + n] - pred_o[:, n])) / scale_w < TOL
    # north-star criterion: per-clip SDR / SIR / SAR within 0.01 dB
    for j in range(B):
        gts = waves[1:, j, :65280].astype(np.float64)
        m_gpu = BSS.get_metrics(got[j, 0], gts, got[j, 1:])
        m_ref = BSS.get_metrics(mix_o[j], gts, pred_o[j])
        for k in ("sdr", "sir", "sar", "si_sdr"):
            assert abs(m_gpu[k] - m_ref[k]) < 0.01, (k, m_gpu[k], m_ref[k])
        assert m_ref["sdr"] > 3.0                                 # the masks do separate
    # host-buffer variant
    wav_h = torch.empty((B, N + 1, 65280)).pin_memory()
    st.eval_host(torch.from_numpy(waves).pin_memory(), torch.from_numpy(vecs).pin_memory(), c(feat), c(scale), c(bias), wav_h)
    assert torch.equal(wav_h.cuda(), wav)
    st.close()


def test_calc_metrics_dropin_matches_reference_golden(golden):
    """calc_metrics (main.py:197-286) with the masks the golden generator fed to the REAL reference function."""
    g = golden("calc_metrics")
    waves = O.synth_sources(2, 2, seed=77)
    mixw = waves.sum(axis=1).astype(np.float32)
    mag, ph = av.stft(torch.from_numpy(mixw).cuda())
    pm = [torch.from_numpy(g[f"pred_mask{n}"].astype(np.float32)).cuda() for n in range(2)]
    args = types.SimpleNamespace(num_mix=2, log_freq=1, stft_frame=1022, stft_hop=256, binary_mask=1, mask_thres=0.5)
    batch = dict(mag_mix=mag[:, None].cpu(), phase_mix=ph[:, None].cpu(), audios=[torch.from_numpy(waves[:, n]) for n in range(2)])
    gm = lambda mix, clean, est, sample_rate=None, metrics_list=None: BSS.get_metrics(mix, clean, est)
    res = av.calc_metrics(batch, dict(pred_masks=pm), args, get_metrics=gm)
    assert res[0] == 0 and len(res) == 8
    assert np.max(np.abs(np.asarray(res[4]) - g["sdr"])) < 0.01
    assert np.max(np.abs(np.asarray(res[5]) - g["sir"])) < 0.01
    assert np.max(np.abs(np.asarray(res[6]) - g["si_sdr"])) < 0.01
    assert abs(res[1] - float(g["sdr_avg"])) < 0.01 and abs(res[3] - float(g["sar_avg"])) < 0.01
    # the batched device evaluator (bss_metrics.bss_eval_batch) on the same clips: no host loop, same numbers
    res_gpu = av.calc_metrics(batch, dict(pred_masks=pm), args, get_metrics="gpu")
    assert np.max(np.abs(np.asarray(res_gpu[4]) - g["sdr"])) < 0.01
    assert np.max(np.abs(np.asarray(res_gpu[5]) - g["sir"])) < 0.01
    assert np.max(np.abs(np.asarray(res_gpu[6]) - g["si_sdr"])) < 0.01
    assert abs(res_gpu[1] - res[1]) < 1e-3 and abs(res_gpu[2] - res[2]) < 1e-3 and abs(res_gpu[3] - res[3]) < 1e-3
    assert abs(res_gpu[7] - res[7]) < 1e-3


def test_n3_and_batch_shapes():
    """Config C3 shape family (N = 3) at a small batch, plus B = 1 (config C1)."""
    for B, N in ((1, 2), (3, 3)):
        waves, feat, vecs, scale, bias = _inputs(B, N, 32, seed=99 + N)
        c = lambda a: torch.from_numpy(a).cuda()
        st = av.Step(B, N, 32, L)
        wav = torch.empty((B, N + 1, 65280), device="cuda")
        st.eval_dev(c(waves), c(vecs), c(feat), c(scale), c(bias), wav, binary_mask=False)
        torch.cuda.synchronize()
        pred = st.buffer("pred_masks").cpu().numpy()
        amp = np.stack([O.stft_magphase(waves[0, b])[0] for b in range(B)])[:, None]
        phase = np.stack([O.stft_magphase(waves[0, b])[1] for b in range(B)])[:, None]
        mix_o, pred_o, _ = O.separate(amp, phase, [pred[n] for n in range(N)], binary_mask=False)
        got = wav.cpu().numpy()
        s = np.max(np.abs(mix_o))
        assert np.max(np.abs(got[:, 0] - mix_o)) / s < TOL
        for n in range(N):
            assert np.max(np.abs(got[:, 1 + n] - pred_o[:, n])) / s < TOL
        st.close()


@pytest.mark.parametrize("N,B", [(2, 2), (3, 3), (4, 1)])
def test_eval_step_from_sources_alone(N, B):
    """Eval step handed the N sources only (avsep_step_set_mix_on_device): the mixture spectrum -- magnitude AND the (re, im) the
    inverse transform consumes -- is the sum of the source transforms (dft_fold3m<N>).  Against the same step with the summed
    waveform transformed as its own signal: spectra equal to fp32 rounding, separated waveforms (ratio masks) within 1e-4."""
    waves, feat, vecs, scale, bias = _inputs(B, N, 32, seed=410 + N)
    c = lambda a: torch.from_numpy(a).cuda()
    src = c(np.ascontiguousarray(waves[1:]))
    st = av.Step(B, N, 32, L)
    st.set_mix_on_device(True)
    outs = []
    for linear in (False, True):
        st.plan.set_mix_linear(linear)
        wav = torch.empty((B, N + 1, 65280), device="cuda")
        st.eval_dev(src, c(vecs), c(feat), c(scale), c(bias), wav, binary_mask=False)
        torch.cuda.synchronize()
        outs.append((st.buffer("mag").clone(), st.buffer("mix_re").clone(), st.buffer("mix_im").clone(), wav))
    (mag0, re0, im0, wav0), (mag1, re1, im1, wav1) = outs
    sc = float(mag0.max())
    assert float((mag1 - mag0).abs().max()) / sc < 2e-5
    assert float((re1 - re0).abs().max()) / sc < 2e-5 and float((im1 - im0).abs().max()) / sc < 2e-5
    assert float((wav1 - wav0).abs().max()) / float(wav0.abs().max()) < 1e-4
    # and against the oracle's STFT of the mixture waveform
    for b in range(B):
        spec = O.stft_complex(waves[0, b])
        got = re1[b].cpu().numpy().astype(np.float64) + 1j * im1[b].cpu().numpy().astype(np.float64)
        assert np.max(np.abs(got - spec)) / np.max(np.abs(spec)) < TOL
    st.close()


def test_prepare_inferdata_caches_per_audio_tensor():
    """inference.py:537-569 calls prepare_inferdata once per video frame on the same audio: the drop-in answers repeat calls on the
    same, unmodified tensor object from its cache (no launch), recomputes after an in-place change or for another tensor."""
    g = torch.Generator(device="cuda").manual_seed(3)
    mag = torch.rand((1, 1, 512, 256), device="cuda", generator=g) + 0.01
    ph = torch.zeros_like(mag)
    av.clear_inferdata_cache()
    _, w0, l0, _ = av.prepare_inferdata((mag, ph), None, None)
    before = av.launch_count()
    _, w1, l1, _ = av.prepare_inferdata((mag, ph), None, None)
    assert av.launch_count() == before and w1 is w0 and l1 is l0
    mag.mul_(2.0)                                                   # in-place change bumps the version counter: recompute
    _, w2, l2, _ = av.prepare_inferdata((mag, ph), None, None)
    assert av.launch_count() > before and w2 is not w0
    torch.cuda.synchronize()
    assert float((w2 - 2.0 * w0).abs().max()) / float(w2.abs().max()) < 1e-6
    _, w3, _, _ = av.prepare_inferdata((mag.clone(), ph), None, None)     # another tensor object: recompute, same values
    _, w4, _, _ = av.prepare_inferdata((mag, ph), None, None, cache=False)
    torch.cuda.synchronize()
    assert w3 is not w2 and torch.equal(w3, w2) and torch.equal(w4, w2)
    av.clear_inferdata_cache()


def test_step_pipeline_equals_serial_steps():
    """StepPipeline (two steps in flight, STFT capped at 80 SMs): every step's results are bit-identical to a lone Step's."""
    B, N, K = 3, 2, 32
    waves, feat, vecs, scale, bias = _inputs(B, N, K, seed=77)
    c = lambda a: torch.from_numpy(a).cuda()
    src, tv, tf, tsc, tb = c(np.ascontiguousarray(waves[1:])), c(vecs), c(feat), c(scale), c(bias)
    st = av.Step(B, N, K, L)
    st.set_mix_on_device(True)
    dfeat0, loss0 = torch.empty_like(tf), torch.empty(1, device="cuda")
    st.train_dev(src, tv, tf, tsc, tb, dfeat0, None, None, None, loss0)
    torch.cuda.synchronize()
    pipe = av.StepPipeline(B, N, K, L, depth=2)
    dfe = [torch.empty_like(tf) for _ in range(2)]
    los = [torch.empty(1, device="cuda") for _ in range(2)]
    for _ in range(6):                                               # direct, capture and replay calls of both slots
        slot = pipe.next()
        with torch.cuda.stream(slot.stream):
            dfe[slot.index].zero_()
            slot.step.train_dev(src, tv, tf, tsc, tb, dfe[slot.index], None, None, None, los[slot.index])
    pipe.synchronize()
    torch.cuda.synchronize()
    for i in range(2):
        assert float(los[i]) == float(loss0) and torch.equal(dfe[i], dfeat0)
    pipe.close()
    st.close()


def test_long_clip_windows_vs_oracle():
    """Config C4 (small): a 3-window clip, fixed masks; every window equals the oracle's single-window path."""
    rng = np.random.default_rng(3)
    L_long = 65535 * 2 + 20000
    src = O.synth_sources(1, 1, L=L_long, seed=11)[0, 0]
    masks_np = rng.random((2, 3, 256, 256)).astype(np.float32)
    seen = {}

    def mask_fn(log_mix, mix_w):
        seen["shape"] = tuple(log_mix.shape)
        return torch.from_numpy(masks_np).cuda()

    out = av.separate_long_clip(torch.from_numpy(src).cuda(), mask_fn, binary_mask=False)
    assert seen["shape"] == (3, 1, 256, 256) and out.shape == (3, 3, 65280)
    wins = av.clip_windows(torch.from_numpy(src)).numpy()          # default stride 65280: seamless tiling
    assert wins.shape[0] == 3
    for i in range(3):
        amp, ph = O.stft_magphase(wins[i])
        mix_o, pred_o, _ = O.separate(amp[None, None], ph[None, None], [masks_np[n][i][None, None] for n in range(2)], binary_mask=False)
        s = max(np.max(np.abs(mix_o)), 1e-9)
        got = out[i].cpu().numpy()
        assert np.max(np.abs(got[0] - mix_o[0])) / s < TOL
        assert np.max(np.abs(got[1:] - pred_o[0])) / s < TOL
    # the mixture channel round-trips over the WHOLE clip (no zero gaps at the window boundaries); only the last
    # samples of the clip, beyond what the zero-padded tail window reconstructs, are outside the check
    full = av.assemble_windows(out[:, 0], 65280, L_long).cpu().numpy()
    covered = min(L_long, 2 * 65280 + 65280)
    assert np.max(np.abs(full[:covered] - src[:covered])) < 1e-4


def test_inference_dropins_vs_oracle(tmp_path):
    """inference.py drop-ins: get_spec (:221-226), prepare_inferdata (:39-52), recover_aud (:433-491)."""
    rng = np.random.default_rng(12)
    src = O.synth_sources(1, 2, seed=21)[0]
    mix = src.sum(axis=0).astype(np.float32)
    amp, phase = av.get_spec(mix)
    assert amp.shape == (1, 1, 512, 256) and phase.shape == (1, 1, 512, 256) and amp.is_cuda
    amp_o, ph_o = O.stft_magphase(mix)
    assert relerr(amp[0, 0].cpu().numpy(), amp_o) < TOL
    args = types.SimpleNamespace(num_mix=2, device="cuda")
    frames, mag_w, log_w, ph = av.prepare_inferdata((amp, phase), "frames-pass-through", args)
    assert frames == "frames-pass-through" and ph is phase
    prep = O.prepare(amp_o[None, None], [], weighted_loss=False)
    assert relerr(mag_w.cpu().numpy(), prep["mag_mix"]) < TOL
    loud = prep["mag_mix"] > 1e-2 * prep["mag_mix"].max()
    assert np.max(np.abs(log_w.cpu().numpy()[loud] - prep["log_mag_mix"][loud])) < 1e-4
    masks = [rng.random((1, 1, 256, 256)).astype(np.float32) for _ in range(2)]
    blk = [np.repeat(np.repeat((rng.random((32, 32)) > 0.5).astype(np.float32) * 0.7 + 0.15, 8, axis=0), 8, axis=1)[None, None] for _ in range(2)]
    meta = dict(pred_masks=[torch.from_numpy(b).cuda() for b in blk])
    written = {}
    pths = av.recover_aud(meta, (amp, phase), str(tmp_path), write=lambda p, w, r: written.__setitem__(os.path.basename(p), (w, r)))
    assert [os.path.basename(p) for p in pths] == ["pred1.wav", "pred2.wav"] and set(written) == {"mix.wav", "pred1.wav", "pred2.wav"}
    assert all(r == 11025 for _, r in written.values())
    mix_o, pred_o, _ = O.separate(amp_o[None, None], ph_o[None, None], blk, log_freq=True, binary_mask=True)
    s0 = np.max(np.abs(mix_o))
    assert np.max(np.abs(written["mix.wav"][0] - mix_o[0])) / s0 < TOL
    for n in range(2):
        assert written[f"pred{n + 1}.wav"][0].dtype == np.float64
        assert np.max(np.abs(written[f"pred{n + 1}.wav"][0] - pred_o[0, n])) / s0 < TOL
        assert meta["pred_masks"][n].shape == (1, 1, 256, 256) and set(np.unique(meta["pred_masks"][n])) <= {0.0, 1.0}


def test_full_size_c3_and_c5_properties():
    """Size-independent properties at BASELINE sizes.  C3: B = 128, N = 3 eval shapes -- with masks that sum to one
    on the linear axis the separated sources add up to the mixture reconstruction (linearity of mask * iSTFT).
    C5: a B = 256 synthesizer batch made of 8 distinct mixtures tiled 32x gives the loss and the tiled gradients
    of the B = 8 batch."""
    B, N = 128, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    mix = (torch.rand((B, L), device="cuda", generator=g) - 0.5) * 0.5
    mag, _, re, im = av.stft(mix, phase=False, complex_parts=True)
    m = torch.rand((N, B, 512, 256), device="cuda", generator=g)
    m = m / m.sum(dim=0, keepdim=True)
    wav = av.separate(m, None, None, re_mix=re, im_mix=im, log_freq=False, binary_mask=False)
    assert wav.shape == (B, N + 1, 65280)
    err = float((wav[:, 1:].sum(dim=1) - wav[:, 0]).abs().max())
    assert err < 2e-5, err
    assert float((wav[:, 0] - mix[:, :65280]).abs().max()) < 2e-5          # and the mixture itself round-trips
    del wav, m, mag, re, im
    K, rep = 32, 32
    feat8 = torch.randn((8, K, 256, 256), device="cuda", generator=g)
    vec8 = torch.sigmoid(torch.randn((8, 2, K), device="cuda", generator=g))
    gt8 = (torch.rand((2, 8, 256, 256), device="cuda", generator=g) > 0.5).float()
    w8 = torch.rand((8, 1, 256, 256), device="cuda", generator=g) + 0.1
    net = av.InnerProd(K).cuda()
    f8 = feat8.clone().requires_grad_(True)
    e8, _ = net.forward_loss(vec8, f8, gt8, w8)
    e8.backward()
    fb = feat8.repeat(rep, 1, 1, 1).requires_grad_(True)
    eb, _ = net.forward_loss(vec8.repeat(rep, 1, 1), fb, gt8.repeat(1, rep, 1, 1), w8.repeat(rep, 1, 1, 1))
    eb.backward()
    assert abs(float(eb.detach()) - float(e8.detach())) < 1e-6 * abs(float(e8.detach()))
    assert float((fb.grad[:8] * rep - f8.grad).abs().max()) <= 1e-6 * float(f8.grad.abs().max())
    assert torch.equal(fb.grad[:8], fb.grad[8 * (rep - 1):])


# ------------------------------------------------------------------------------------------------------------------
# Full-size comparators: the whole step at the BASELINE.json sizes against the reference's own sequence of torch calls
# (oracle/torch_ref_path.py) executed by stock PyTorch kernels on the same GPU.
# ------------------------------------------------------------------------------------------------------------------
def _full_inputs(B, N, K=32, seed=4321):
    import bench
    waves = bench.make_waves(B, N, seed).cuda()                   # [N+1, B, L], index 0 = mixture
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    feat = torch.randn((B, K, 256, 256), device="cuda", generator=g)
    vecs = torch.sigmoid(torch.randn((B, N, K), device="cuda", generator=g))
    scale = 1 + 0.05 * torch.randn(K, device="cuda", generator=g)
    bias = torch.full((1,), 0.02, device="cuda")
    return waves, feat, vecs, scale, bias


def test_full_size_train_step_vs_stock_torch():
    """Config C2 at full size (B = 80, N = 2): loss, masks, targets and all four gradients of the fused step against
    the reference's torch call sequence on the GPU.  Tolerance: max-norm relative 1e-4 (north star); binary targets by
    mismatch fraction (an exact tie of mag_n vs 0.5 mag_mix can flip at fp32 rounding level)."""
    from oracle import torch_ref_path as TP
    B, N, K = 80, 2, 32
    waves, feat, vecs, scale, bias = _full_inputs(B, N, K)
    st = av.Step(B, N, K, L)
    dfeat = torch.empty_like(feat)
    dvecs = torch.empty_like(vecs)
    dscale = torch.empty(K, device="cuda")
    dbias = torch.empty(1, device="cuda")
    loss = torch.empty(1, device="cuda")
    st.train_dev(waves, vecs, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss)      # explicit mixture waveform: 3 transforms
    torch.cuda.synchronize()
    gt_ours = st.buffer("gt_masks")                               # [N, B, 1, 256, 256]
    pred_ours = st.buffer("pred_masks")
    w_ours = st.buffer("weights")
    # reference sequence (stock kernels)
    f = feat.clone().requires_grad_(True)
    v = vecs.clone().requires_grad_(True)
    sc = scale.clone().requires_grad_(True)
    bi = bias.clone().requires_grad_(True)
    wb = waves.permute(1, 0, 2).contiguous()                       # [B, N+1, L]
    amp, _ = TP.stft_magphase(wb.reshape(B * (N + 1), L))
    amp = amp.view(B, N + 1, 1, 512, 256)
    _, _, _, gt_ref, w_ref = TP.prepare(amp[:, 0], [amp[:, 1 + n] for n in range(N)])
    preds = [torch.sigmoid(TP.innerprod(v[:, n, :], f, sc, bi)) for n in range(N)]
    err = TP.bce_list(preds, gt_ref, w_ref)
    err.backward()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(w_ours, w_ref) < TOL
    # ties of mag_n against 0.5 mag_mix flip where the bin sits at the noise floor (a 5e-6 max-norm error is a large LOCAL
    # relative error there); such pixels carry next to no loss weight, so the weighted fraction is the tight criterion
    mism = sum(float((gt_ours[n] != gt_ref[n]).float().mean()) for n in range(N)) / N
    wmism = sum(float(((gt_ours[n] != gt_ref[n]).float() * w_ref).sum() / w_ref.sum()) for n in range(N)) / N
    assert mism < 3e-4 and wmism < 3e-5, (mism, wmism)
    for n in range(N):
        assert rel(pred_ours[n], preds[n].detach()) < TOL
    # the loss / gradients of OUR targets vs the reference's targets differ by the flipped pixels only (<= 2e-5 of them,
    # each worth up to ~10 / numel): compare with the reference re-evaluated on our targets for a tight bound as well
    assert abs(float(loss) - float(err)) / float(err) < 2e-4
    f2 = feat.clone().requires_grad_(True)
    v2 = vecs.clone().requires_grad_(True)
    sc2 = scale.clone().requires_grad_(True)
    bi2 = bias.clone().requires_grad_(True)
    preds2 = [torch.sigmoid(TP.innerprod(v2[:, n, :], f2, sc2, bi2)) for n in range(N)]
    err2 = TP.bce_list(preds2, [gt_ours[n] for n in range(N)], w_ours)
    err2.backward()
    assert abs(float(loss) - float(err2)) / float(err2) < TOL
    assert rel(dfeat, f2.grad) < TOL
    assert rel(dvecs, v2.grad) < TOL
    assert rel(dscale, sc2.grad) < TOL
    assert abs(float(dbias) - float(bi2.grad)) / abs(float(bi2.grad)) < TOL
    # the default training form (sources only, mixture spectrum = X0 + X1) agrees with the explicit-mixture step
    st.set_mix_on_device(True)
    loss_m = torch.empty(1, device="cuda")
    st.train_dev(waves[1:], vecs, feat, scale, bias, dfeat, dvecs, dscale, dbias, loss_m)
    torch.cuda.synchronize()
    assert abs(float(loss_m) - float(loss)) / float(loss) < 1e-3
    assert float((st.buffer("gt_masks") != gt_ours).float().mean()) < 1e-4
    st.close()


def test_full_size_eval_step_vs_stock_torch():
    """Config C3 at full size (B = 128, N = 3): the separated waveforms of the eval step against the reference's torch
    call sequence (fp64 iSTFT) on the GPU.  Ratio masks for the waveform comparison (max-norm relative 1e-4); the
    thresholded un-warped masks of the binary setting are compared by mismatch fraction."""
    from oracle import torch_ref_path as TP
    import torch.nn.functional as F
    B, N, K = 128, 3, 32
    waves, feat, vecs, scale, bias = _full_inputs(B, N, K, seed=977)
    st = av.Step(B, N, K, L)
    wav = torch.empty((B, N + 1, 65280), device="cuda")
    st.eval_dev(waves, vecs, feat, scale, bias, wav, binary_mask=False)
    torch.cuda.synchronize()
    ref = TP.eval_step(waves.permute(1, 0, 2).contiguous(), feat, vecs, scale, bias, binary_mask=False)    # [B, N+1, 65280] float64
    scale_ref = float(ref[:, 0].abs().max())
    err = float((wav.double() - ref).abs().max()) / scale_ref
    assert err < TOL, err
    # binary setting: our un-warp + threshold against grid_sample + threshold on the same predicted masks
    pred = st.buffer("pred_masks")                                 # [N, B, 1, 256, 256]
    grid = torch.from_numpy(TP.warpgrid(B, 512, 256, warp=False)).cuda()
    mism = 0.0
    for n in range(N):
        lin_ref = (F.grid_sample(pred[n], grid, align_corners=False) > 0.5)
        lin_ours = av.warp(pred[n], 512, warp=False) > 0.5
        mism += float((lin_ref != lin_ours).float().mean())
    assert mism / N < 1e-5, mism / N
    # and the binary-mask waveforms: every clip except those hit by a flipped pixel agrees to the tolerance
    st.eval_dev(waves, vecs, feat, scale, bias, wav, binary_mask=True)
    torch.cuda.synchronize()
    refb = TP.eval_step(waves.permute(1, 0, 2).contiguous(), feat, vecs, scale, bias, binary_mask=True)
    per_clip = (wav.double() - refb).abs().amax(dim=(1, 2)) / scale_ref
    assert float(per_clip.median()) < TOL and float((per_clip > TOL).float().mean()) <= 0.05, per_clip
    st.close()
