"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, its host-side tap tables equal the reference's warpgrid, and the product path
refuses to run (loudly) without an sm_100 GPU -- there is no CPU fallback to fall into."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import np_oracle as O

import avsep_b200 as av

HEADER = os.path.join(ROOT, "include", "avsep_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avsep_[a-z0-9_]+)\s*\(", src)))


def test_library_is_in_tree_and_loads():
    assert os.path.dirname(av.LIB_PATH).startswith(ROOT)
    assert os.path.exists(av.LIB_PATH)
    assert av._lib.lib.avsep_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 30
    raw = C.CDLL(av.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/avsep_b200.h but not exported"
        assert n in av._lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    for n in av._lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_tap_tables_equal_reference_warpgrid(golden):
    g = golden("warpgrid")
    lib = av._lib.lib
    for kind, n_in, n_out, coords in ((0, 512, 256, g["warp_y"]), (1, 256, 512, g["unwarp_y"]), (2, 256, 256, g["warp_x"])):
        i0 = np.zeros(n_out, dtype=np.int32)
        fr = np.zeros(n_out, dtype=np.float32)
        assert lib.avsep_warp_taps_host(kind, n_in, n_out, i0.ctypes.data, fr.ctypes.data) == 0
        oi, of = O.axis_taps(n_out, n_in, coords)
        assert np.array_equal(i0, oi) and np.array_equal(fr, of)
    # taps described in SURVEY.md section 8(a) row 3
    i0 = np.zeros(256, dtype=np.int32); fr = np.zeros(256, dtype=np.float32)
    lib.avsep_warp_taps_host(0, 512, 256, i0.ctypes.data, fr.ctypes.data)
    assert list(i0[:4]) == [-1, -1, 0, 0] and list(i0[-3:]) == [498, 505, 511]
    assert np.max(np.diff(i0)) == 7


def test_bad_arguments_are_reported():
    lib = av._lib.lib
    assert lib.avsep_warp_taps_host(7, 4, 4, None, None) != 0
    assert "bad arguments" in av._lib.last_error()
    h = C.c_void_p()
    assert lib.avsep_plan_create(C.byref(h), 0, 65535) == 1          # AVSEP_ERR_INVALID
    assert lib.avsep_plan_create(C.byref(h), 4, 100) == 1
    assert "too short" in av._lib.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(av.AvsepError, match="no CPU"):
        av.device_check()
    with pytest.raises(av.AvsepError, match="CUDA tensor"):
        av.stft(torch.zeros(2, 65535))
    with pytest.raises(av.AvsepError, match="CUDA tensor"):
        av.InnerProd(32)(torch.zeros(2, 32), torch.zeros(2, 32, 8, 8))
    with pytest.raises(av.AvsepError, match="CUDA tensor"):
        av.BCELoss()([torch.rand(4)], [torch.rand(4)], torch.ones(4))
    h = C.c_void_p()
    assert av._lib.lib.avsep_plan_create(C.byref(h), 4, 65535) == 3   # AVSEP_ERR_NO_DEVICE
    assert av._lib.lib.avsep_step_create(C.byref(h), 1, 2, 32, 65535) == 3


def test_product_does_not_import_the_oracle():
    pkg = os.path.dirname(av.LIB_PATH)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+[\w.]*oracle", txt, re.M), f"{f} imports the oracle"
                assert "reference" not in [m for m in re.findall(r"^\s*(?:from|import)\s+([\w.]+)", txt, re.M)]


def test_module_state_dict_matches_reference_keys():
    m = av.InnerProd(32)
    assert list(m.state_dict().keys()) == ["scale", "bias"]          # synthesizer_net.py:9-10
    assert m.scale.shape == (32,) and m.bias.shape == (1,)
    assert list(av.Bias().state_dict().keys()) == ["bias"]
    with pytest.raises(Exception, match="Unkown activation"):
        av.activate(torch.zeros(1), "swish")                          # models/__init__.py:28


def test_warpgrid_compat(golden):
    g = golden("warpgrid")
    gw = av.warpgrid(2, 256, 256, warp=True)
    assert gw.shape == (2, 256, 256, 2) and gw.dtype == np.float32
    assert np.array_equal(gw[1, :, 5, 1], g["warp_y"]) and np.array_equal(gw[0, 3, :, 0], g["warp_x"])


def test_clip_windows_and_assemble_roundtrip():
    """Long-clip helper (config C4): 60 s @ 11025 Hz -> windows.  The default stride is what a window reconstructs
    (65280 samples), so the assembled clip has no gaps; stride 32768 gives the overlapping secondary setting."""
    L = 661500
    x = torch.arange(L, dtype=torch.float32)
    w = av.clip_windows(x)
    assert w.shape == (11, 65535) and float(w[10, -1]) == 0.0 and float(w[1, 0]) == 65280.0
    assert av.clip_windows(x, 65535, 32768).shape[0] == 20
    assert av.clip_windows(x[:1000]).shape == (1, 65535)
    # reconstruction windows are 65280 long and tile the clip seamlessly: gap-free round trip over the whole clip
    rec = w[:, :65280]
    y = av.assemble_windows(rec, 65280, L)
    assert torch.equal(y, x)
    # an explicit stride of 65535 leaves the 255 samples no window reconstructs at zero
    w_gap = av.clip_windows(x, 65535, 65535)
    y_gap = av.assemble_windows(w_gap[:, :65280], 65535, L)
    assert torch.equal(y_gap[:65280], x[:65280]) and float(y_gap[65280:65535].abs().max()) == 0.0
    # overlapping stride: later windows take over at their offset
    w2 = av.clip_windows(x, 65535, 32768)
    y2 = av.assemble_windows(w2[:, :65280], 32768, L)
    assert torch.equal(y2[:L - 300], x[:L - 300])


def test_bench_reads_ncu_traffic_from_committed_profile():
    """bench.py fills roofline.traffic from profiles/r1_ncu_full_summary.txt (dram read + write of one launch)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.ncu_traffic("synth_train_kernel")
    assert t is not None and 1.0e9 < t < 2.0e9            # 1.34 GB of feat / dfeat + masks, targets, weights
    # every kernel name bench.py's roofline blocks may look up is present in the committed ncu summary
    for name in ("dft_fold3m_fwd_kernel<2, 2>", "dft_fold3m_fwd_kernel<2, 3>", "dft_fold3_fwd_kernel", "prepare_bulk_kernel",
                 "synth_fwd_staged_kernel", "spec_prep_tma_kernel", "idft_fold2_kernel"):
        assert bench.ncu_traffic(name) is not None, name
    assert bench.ncu_traffic("no_such_kernel") is None
