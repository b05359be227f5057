"""Generate tests/golden/*.npz by running the REAL reference code from /root/reference.

TEST INFRASTRUCTURE ONLY.  Run here (the build container), never on the GPU box:

    python oracle/make_golden.py            # writes tests/golden/*.npz and self-checks the oracle

What is executed from the reference, unmodified:
  * utils.warpgrid                      (utils.py:12-26)      -- module imported with a stub `librosa`
  * models/synthesizer_net.py           InnerProd, Bias        -- module imported by file path
  * models/criterion.py                 BCELoss, L1Loss, L2Loss, PitWrapper
  * main.py NetWrapper.prepare          (main.py:51-95)        -- function source extracted with `ast`
    (main.py itself cannot be imported: tkinter / asteroid / imageio / soundfile / mmcv are absent)
  * main.py calc_metrics                (main.py:197-286)      -- same extraction; its two un-vendored
    callees are stubbed: istft_reconstruction -> torch.istft based stand-in, get_metrics -> oracle/bss_eval
  * torch.nn.functional.grid_sample / binary_cross_entropy / torch.stft / torch.istft: the library
    calls the reference makes (torch.stft/istft are the stand-in for the absent librosa).
Inputs are regenerated from seeds by the tests (numpy PCG64 is platform-stable), so only outputs are stored.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import np_oracle as O          # noqa: E402
from oracle import bss_eval as BSS         # noqa: E402
from oracle import torch_ref_path as TP    # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    ref_utils = _load("ref_utils", os.path.join(REF, "utils.py"))
    ref_synth = _load("ref_synth", os.path.join(REF, "models", "synthesizer_net.py"))
    ref_crit = _load("ref_crit", os.path.join(REF, "models", "criterion.py"))
    return ref_utils, ref_synth, ref_crit


def extract_function(path, qualname):
    """Return the source of a (possibly nested-in-class) function of a file we cannot import."""
    src = open(path).read()
    tree = ast.parse(src)
    parts = qualname.split(".")
    node = tree
    for p in parts:
        for child in ast.iter_child_nodes(node):
            if isinstance(child, (ast.FunctionDef, ast.ClassDef)) and child.name == p:
                node = child
                break
        else:
            raise KeyError(qualname)
    seg = ast.get_source_segment(src, node)
    import textwrap
    return textwrap.dedent(seg)


def rng_inputs(seed, *shapes):
    g = np.random.default_rng(seed)
    return [g.standard_normal(s).astype(np.float32) for s in shapes]


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_utils, ref_synth, ref_crit = load_reference()
    report = {}

    # ---- 1. warpgrid ------------------------------------------------------------------
    gw = ref_utils.warpgrid(2, 256, 256, warp=True)
    gu = ref_utils.warpgrid(2, 512, 256, warp=False)
    assert np.array_equal(gw, O.warpgrid(2, 256, 256, True))
    assert np.array_equal(gu, O.warpgrid(2, 512, 256, False))
    assert np.array_equal(gw, TP.warpgrid(2, 256, 256, True))
    np.savez_compressed(os.path.join(OUT, "warpgrid.npz"),
                        warp_y=gw[0, :, 0, 1], warp_x=gw[0, 0, :, 0],
                        unwarp_y=gu[0, :, 0, 1], unwarp_x=gu[0, 0, :, 0])

    # ---- 2. grid_sample with the reference grid ------------------------------------------
    (xw,) = rng_inputs(11, (2, 1, 512, 256))
    xw = np.abs(xw)
    yw = F.grid_sample(torch.from_numpy(xw), torch.from_numpy(gw), align_corners=False).numpy()
    (xu,) = rng_inputs(12, (2, 1, 256, 256))
    yu = F.grid_sample(torch.from_numpy(xu), torch.from_numpy(gu), align_corners=False).numpy()
    report["grid_sample_warp"] = relerr(O.grid_sample(xw, gw), yw)
    report["grid_sample_unwarp"] = relerr(O.grid_sample(xu, gu), yu)
    # narrow-T case (T=24) stored whole, full-T case stored for batch 1 only
    (xs,) = rng_inputs(13, (1, 1, 512, 24))
    gs = ref_utils.warpgrid(1, 256, 24, warp=True)
    ys = F.grid_sample(torch.from_numpy(xs), torch.from_numpy(gs), align_corners=False).numpy()
    report["grid_sample_T24"] = relerr(O.grid_sample(xs, gs), ys)
    np.savez_compressed(os.path.join(OUT, "grid_sample.npz"),
                        warp_out_b1=yw[1], unwarp_out_b1=yu[1], warp_T24=ys)

    # ---- 3. prepare (real main.py source) ---------------------------------------------------
    src = extract_function(os.path.join(REF, "main.py"), "NetWrapper.prepare")
    ns = dict(torch=torch, F=F, warpgrid=ref_utils.warpgrid, np=np)
    exec(compile(src, "main.py:NetWrapper.prepare", "exec"), ns)
    ref_prepare = ns["prepare"]
    prep_out = {}
    for tag, (binary, weighted) in {"bin_w": (1, 1), "ratio_now": (0, 0)}.items():
        g = np.random.default_rng(21)
        mags = [np.abs(g.standard_normal((2, 1, 512, 64))).astype(np.float32) * 3 for _ in range(2)]
        mag_mix = (mags[0] + mags[1] * np.float32(0.7)).astype(np.float32)
        args = types.SimpleNamespace(num_mix=2, log_freq=1, device="cpu", weighted_loss=weighted,
                                     binary_mask=binary, fusion_type="none")
        batch = dict(mag_mix=torch.from_numpy(mag_mix.copy()), mags=[torch.from_numpy(m.copy()) for m in mags])
        self_stub = types.SimpleNamespace(load_clips=False)
        r_mags, r_mix, r_log, r_gt, r_w = ref_prepare(self_stub, batch, args, False, False)
        o = O.prepare(mag_mix, mags, log_freq=True, weighted_loss=bool(weighted), binary_mask=bool(binary))
        t_mags, t_mix, t_log, t_gt, t_w = TP.prepare(torch.from_numpy(mag_mix.copy()),
                                                     [torch.from_numpy(m.copy()) for m in mags],
                                                     True, bool(weighted), bool(binary))
        assert torch.equal(t_mix, r_mix) and torch.equal(t_w, r_w) and torch.equal(t_gt[1], r_gt[1])
        report[f"prepare_{tag}_mix"] = relerr(o["mag_mix"], r_mix.numpy())
        report[f"prepare_{tag}_log"] = relerr(o["log_mag_mix"], r_log.numpy())
        report[f"prepare_{tag}_w"] = relerr(o["weights"], r_w.numpy())
        if binary:
            report[f"prepare_{tag}_gt_mismatch_frac"] = float(np.mean(o["gt_masks"][0] != r_gt[0].numpy()))
        else:
            report[f"prepare_{tag}_gt"] = relerr(o["gt_masks"][0], r_gt[0].numpy())
        prep_out.update({f"{tag}_mag_mix": r_mix.numpy(), f"{tag}_mags0": r_mags[0].numpy(),
                         f"{tag}_mags1": r_mags[1].numpy(), f"{tag}_log": r_log.numpy(),
                         f"{tag}_w": r_w.numpy(), f"{tag}_gt0": r_gt[0].numpy(), f"{tag}_gt1": r_gt[1].numpy()})
    np.savez_compressed(os.path.join(OUT, "prepare.npz"), **prep_out)

    # ---- 4. synthesizer + criterion (real modules, autograd grads) --------------------------------
    B, K, H, W, N = 2, 32, 16, 24, 2
    g = np.random.default_rng(31)
    feat = g.standard_normal((B, K, H, W)).astype(np.float32)
    vecs = (1 / (1 + np.exp(-g.standard_normal((B, N, K))))).astype(np.float32)
    scale = (1 + 0.1 * g.standard_normal(K)).astype(np.float32)
    bias = np.array([0.05], dtype=np.float32)
    gts = [(g.random((B, 1, H, W)) > 0.5).astype(np.float32) for _ in range(N)]
    wgt = np.clip(np.abs(g.standard_normal((B, 1, H, W))), 1e-3, 10).astype(np.float32)
    syn = {}
    net = ref_synth.InnerProd(K)
    with torch.no_grad():
        net.scale.copy_(torch.from_numpy(scale))
        net.bias.copy_(torch.from_numpy(bias))
    for kind, cls in (("bce", ref_crit.BCELoss), ("l1", ref_crit.L1Loss), ("l2", ref_crit.L2Loss)):
        tf = torch.from_numpy(feat).requires_grad_(True)
        tv = torch.from_numpy(vecs).requires_grad_(True)
        net.zero_grad()
        preds = [torch.sigmoid(net(tv[:, n, :], tf)) for n in range(N)]
        err = cls()(preds, [torch.from_numpy(t) for t in gts], torch.from_numpy(wgt))
        err.backward()
        o = O.synth_loss_and_grads(kind, feat, vecs, scale, bias, gts, wgt)
        report[f"synth_{kind}_loss"] = abs(o["loss"] - float(err)) / abs(float(err))
        report[f"synth_{kind}_dfeat"] = relerr(o["dfeat"], tf.grad.numpy())
        report[f"synth_{kind}_dvecs"] = relerr(o["dvecs"], tv.grad.numpy())
        report[f"synth_{kind}_dscale"] = relerr(o["dscale"], net.scale.grad.numpy())
        report[f"synth_{kind}_dbias"] = abs(o["dbias"] - float(net.bias.grad)) / abs(float(net.bias.grad))
        syn.update({f"{kind}_loss": np.float64(float(err)), f"{kind}_dfeat": tf.grad.numpy(),
                    f"{kind}_dvecs": tv.grad.numpy(), f"{kind}_dscale": net.scale.grad.numpy().copy(),
                    f"{kind}_dbias": net.bias.grad.numpy().copy()})
        if kind == "bce":
            syn["pred0"] = preds[0].detach().numpy()
            syn["pred1"] = preds[1].detach().numpy()
    with torch.no_grad():
        z = net(torch.from_numpy(vecs[:, 0]), torch.from_numpy(feat)).numpy()
        zn = net.forward_nosum(torch.from_numpy(vecs[:, 0]), torch.from_numpy(feat)).numpy()
        fi = g.standard_normal((B, K, 3, 5)).astype(np.float32)
        zp = net.forward_pixelwise(torch.from_numpy(fi), torch.from_numpy(feat)).numpy()
        bnet = ref_synth.Bias()
        bnet.bias.copy_(torch.from_numpy(bias))
        zb = bnet(torch.from_numpy(vecs[:, 1]), torch.from_numpy(feat)).numpy()
    report["innerprod_fwd"] = relerr(O.innerprod_forward(vecs[:, 0], feat, scale, bias), z)
    report["innerprod_nosum"] = relerr(O.innerprod_forward_nosum(vecs[:, 0], feat, scale, bias), zn)
    report["innerprod_pixelwise"] = relerr(O.innerprod_forward_pixelwise(fi, feat, scale, bias), zp)
    report["bias_fwd"] = relerr(O.innerprod_forward(vecs[:, 1], feat, None, bias), zb)
    syn.update(z=z, z_nosum=zn, z_pixelwise=zp, z_bias=zb)
    # saturation semantics: z = +-40 rows  
    ps = torch.tensor([40.0, -40.0, 17.0, -104.0, 0.3]).requires_grad_(True)
    ts = torch.tensor([0.0, 1.0, 0.0, 1.0, 1.0])
    ls = F.binary_cross_entropy(torch.sigmoid(ps), ts, weight=torch.full((5,), 2.0), reduction="sum")
    ls.backward()
    syn.update(sat_z=ps.detach().numpy(), sat_t=ts.numpy(), sat_loss=np.float64(float(ls)), sat_grad=ps.grad.numpy())
    np.savez_compressed(os.path.join(OUT, "synth.npz"), **syn)

    # ---- 5. PIT --------------------------------------------------------------------------------
    pitw = ref_crit.PitWrapper(F.binary_cross_entropy)
    pit_out = {}
    for S in (2, 3):
        g = np.random.default_rng(40 + S)
        pr = g.random((3, 12, 10, S)).astype(np.float32) * 0.98 + 0.01
        tg = (g.random((3, 12, 10, S)) > 0.5).astype(np.float32)
        # make sample 1 a permuted copy so the optimum is non-identity
        pr[1] = np.clip(tg[1][..., ::-1] * 0.9 + 0.05, 0.01, 0.99)
        ww = np.repeat(np.abs(g.standard_normal((3, 12, 10, 1))).astype(np.float32), S, axis=-1)
        tpr = torch.from_numpy(pr).requires_grad_(True)
        loss, perms = pitw(tpr, torch.from_numpy(tg), torch.from_numpy(ww))
        (loss * torch.tensor([1.0, 2.0, 0.5])).sum().backward()
        ol, op = O.pit("bce", pr, tg, ww)
        assert [tuple(p) for p in perms] == [tuple(p) for p in op], (perms, op)
        report[f"pit_S{S}"] = relerr(ol, loss.detach().numpy())
        re = pitw.reorder_tensor(torch.from_numpy(pr), perms).numpy()
        assert np.array_equal(re, O.pit_reorder(pr, op))
        og = O.pit_grad("bce", pr, tg, ww, op, np.array([1.0, 2.0, 0.5]))
        report[f"pit_S{S}_grad"] = relerr(og, tpr.grad.numpy())
        pit_out.update({f"S{S}_loss": loss.detach().numpy(), f"S{S}_perms": np.asarray(perms, dtype=np.int64),
                        f"S{S}_grad": tpr.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "pit.npz"), **pit_out)

    # ---- 6. STFT / iSTFT (torch.stft = librosa stand-in) ----------------------------------------
    src_w = O.synth_sources(1, 2, seed=1234)[0]                 # [2, 65535]
    mix_w = (src_w[0] + src_w[1]).astype(np.float32)
    amp_t, ph_t = TP.stft_magphase(torch.from_numpy(np.stack([mix_w, src_w[0]])))
    amp_o, ph_o = O.stft_magphase(mix_w)
    report["stft_amp"] = relerr(amp_o, amp_t[0].numpy())
    spec_t = (amp_t[0].double() * torch.exp(1j * ph_t[0].double())).numpy()
    spec_o = amp_o.astype(np.float64) * np.exp(1j * ph_o.astype(np.float64))
    report["stft_complex"] = relerr(np.abs(spec_o - spec_t), np.abs(spec_t)) if False else \
        float(np.max(np.abs(spec_o - spec_t)) / np.max(np.abs(spec_t)))
    short = src_w[1][:3000].copy()
    amp_s, ph_s = TP.stft_magphase(torch.from_numpy(short[None]))
    report["stft_short_amp"] = relerr(O.stft_magphase(short)[0], amp_s[0].numpy())
    g = np.random.default_rng(51)
    mask = (g.random((512, 256)) > 0.4).astype(np.float32)
    wav_t = TP.istft_reconstruction(amp_t[0:1] * torch.from_numpy(mask)[None], ph_t[0:1])[0].numpy()
    wav_o = O.istft_reconstruction(amp_t[0].numpy() * mask, ph_t[0].numpy())
    report["istft"] = relerr(wav_o, wav_t)
    rt = TP.istft_reconstruction(amp_t[0:1], ph_t[0:1])[0].numpy()
    report["stft_istft_roundtrip_vs_wave"] = float(np.max(np.abs(rt - mix_w[:65280])))
    frames = np.array([0, 1, 2, 3, 127, 128, 252, 253, 254, 255])
    np.savez_compressed(os.path.join(OUT, "stft.npz"), frames=frames,
                        amp_mix=amp_t[0].numpy()[:, frames], phase_mix=ph_t[0].numpy()[:, frames],
                        amp_src0=amp_t[1].numpy()[:, frames],
                        amp_short=amp_s[0].numpy(), phase_short=ph_s[0].numpy(),
                        istft_masked_head=wav_t[:4096], istft_masked_mid=wav_t[30000:34096],
                        istft_masked_tail=wav_t[-4096:], istft_len=np.int64(wav_t.shape[0]))

    # ---- 7. calc_metrics control flow (real main.py source, stubbed third-party callees) ---------
    src = extract_function(os.path.join(REF, "main.py"), "calc_metrics")
    def _istft_stub(mag, phase, hop_length=256):
        return TP.istft_reconstruction(torch.from_numpy(mag)[None], torch.from_numpy(phase)[None])[0].numpy()
    def _get_metrics_stub(mix, clean, estimate, sample_rate=None, metrics_list=None):
        return BSS.get_metrics(mix, clean, estimate)
    ns = dict(torch=torch, F=F, np=np, warpgrid=ref_utils.warpgrid, AverageMeter=ref_utils.AverageMeter,
              istft_reconstruction=_istft_stub, get_metrics=_get_metrics_stub)
    exec(compile(src, "main.py:calc_metrics", "exec"), ns)
    ref_calc = ns["calc_metrics"]
    waves = O.synth_sources(2, 2, seed=77)                       # [2,2,L]
    mixw = waves.sum(axis=1).astype(np.float32)
    amp_m, ph_m = TP.stft_magphase(torch.from_numpy(mixw))
    amps = [TP.stft_magphase(torch.from_numpy(waves[:, n]))[0] for n in range(2)]
    _, mm_w, _, gt, _ = TP.prepare(amp_m[:, None], [a[:, None] for a in amps])
    g = np.random.default_rng(61)
    pm = [torch.sigmoid((gt[n] * 2 - 1) * 2 + torch.from_numpy(g.standard_normal(gt[n].shape).astype(np.float32)))
          for n in range(2)]
    args = types.SimpleNamespace(num_mix=2, log_freq=1, device="cpu", stft_frame=1022, stft_hop=256,
                                 binary_mask=1, mask_thres=0.5)
    batch = dict(mag_mix=amp_m[:, None], phase_mix=ph_m[:, None],
                 audios=[torch.from_numpy(waves[:, n]) for n in range(2)])
    res = ref_calc(batch, dict(pred_masks=pm), args)
    mix_o, pred_o, _ = O.separate(amp_m[:, None].numpy(), ph_m[:, None].numpy(), [p.numpy() for p in pm])
    mets = [BSS.get_metrics(mix_o[j], waves[j][:, :mix_o.shape[1]], pred_o[j]) for j in range(2)]
    report["calc_metrics_sdr_abs_db"] = float(max(abs(m["sdr"] - r) for m, r in zip(mets, res[4])))
    report["calc_metrics_sir_abs_db"] = float(max(abs(m["sir"] - r) for m, r in zip(mets, res[5])))
    np.savez_compressed(os.path.join(OUT, "calc_metrics.npz"),
                        pred_mask0=pm[0].numpy().astype(np.float16), pred_mask1=pm[1].numpy().astype(np.float16),
                        sdr=np.asarray(res[4]), sir=np.asarray(res[5]), si_sdr=np.asarray(res[6]),
                        sdr_avg=np.float64(res[1]), sir_avg=np.float64(res[2]), sar_avg=np.float64(res[3]))

    width = max(len(k) for k in report)
    for k, v in report.items():
        print(f"{k:<{width}}  {v:.3e}")
    bad = {k: v for k, v in report.items() if v > 2e-5 and "mismatch" not in k and "roundtrip" not in k}
    assert not bad, bad
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
