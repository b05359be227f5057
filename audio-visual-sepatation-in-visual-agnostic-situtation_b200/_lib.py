"""ctypes binding of libavsep_b200.so (the C-ABI declared in include/avsep_b200.h).

There is no CPU or PyTorch fallback behind these symbols: if the library cannot be loaded the
import fails, and on a machine without an sm_100 GPU every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libavsep_b200.so")

c_float_p = C.POINTER(C.c_float)
c_void_pp = C.POINTER(C.c_void_p)


class AvsepError(RuntimeError):
    """Raised when a C-ABI call returns non-zero (mirrors torch's RuntimeError convention)."""


def _load() -> C.CDLL:
    """Load the in-tree library, never a stale one: with nvcc present build() is consulted on every import (it returns
    immediately when the source fingerprint matches); without nvcc (a GPU box that received the prebuilt .so) the
    recorded fingerprint must equal the fingerprint of the sources that travelled with it."""
    from . import build as _build
    try:
        _build._nvcc()
        have_nvcc = True
    except RuntimeError:
        have_nvcc = False
    if have_nvcc:
        _build.build()
    else:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing and nvcc is not available to build it (avsep_b200 has no CPU fallback)")
        stamp = os.path.join(_build.OBJ_DIR, "fingerprint")
        if not os.path.exists(stamp) or open(stamp).read() != _build._fingerprint():
            raise ImportError(f"{LIB_PATH} was not built from the sources next to it (fingerprint mismatch) and nvcc is "
                              "not available to rebuild it")
    return C.CDLL(LIB_PATH)


lib = _load()

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/avsep_b200.h declares
SIGNATURES = {
    "avsep_abi_version": (_I, []),
    "avsep_last_error": (C.c_char_p, []),
    "avsep_device_check": (_I, [_I]),
    "avsep_launch_count": (C.c_uint64, []),
    "avsep_plan_create": (_I, [c_void_pp, _I, _I]),
    "avsep_plan_destroy": (_I, [_P]),
    "avsep_plan_frames": (_I, [_P]),
    "avsep_plan_istft_len": (_I, [_P]),
    "avsep_plan_workspace_bytes": (_SZ, [_P]),
    "avsep_plan_set_fold": (_I, [_P, _I]),
    "avsep_plan_set_inverse_fold": (_I, [_P, _I]),
    "avsep_plan_set_mix_linear": (_I, [_P, _I]),
    "avsep_plan_set_sm_budget": (_I, [_P, _I]),
    "avsep_plan_set_spec_staged": (_I, [_P, _I]),
    "avsep_stft_mix": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "avsep_stft_mix_reim": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "avsep_stft": (_I, [_P, _P, _I, _P, _P, _P, _P, _P]),
    "avsep_stft_reim": (_I, [_P, _P, _I, _P, _P, _P, _I, _P]),
    "avsep_separate": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _F, _I, _P, _P]),
    "avsep_istft": (_I, [_P, _P, _P, _I, _P, _P]),
    "avsep_debug_raw_gemm": (_I, [_P, _I, _I, _P, _P]),
    "avsep_debug_gemm_only": (_I, [_P, _I, _I, _P, _P]),
    "avsep_debug_basis": (_I, [_P, _I, _P]),
    "avsep_debug_staged_frames": (_I, [_P, _I, _P]),
    "avsep_warp_taps_host": (_I, [_I, _I, _I, _P, _P]),
    "avsep_warp": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "avsep_prepare": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "avsep_synth_scratch_bytes": (_SZ, [_I, _I, _I, _I]),
    "avsep_synth_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "avsep_synth_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_synth_train": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_synth_nosum": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "avsep_synth_pixelwise": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "avsep_loss_scratch_bytes": (_SZ, [_I, _SZ]),
    "avsep_loss_forward": (_I, [_I, _P, _P, _P, _I, _SZ, _P, _P, _P]),
    "avsep_loss_backward": (_I, [_I, _P, _P, _P, _I, _SZ, _P, _P, _P]),
    "avsep_scale_if_needed": (_I, [_P, _SZ, _P, _P]),
    "avsep_pit_lossmat": (_I, [_I, _P, _P, _P, _I, _SZ, _I, _P, _P]),
    "avsep_synth_forward_pit": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "avsep_plane_gather": (_I, [_P, _P, _P, _F, _I, _I, _SZ, _P, _P]),
    "avsep_pit_search": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "avsep_pit_reorder": (_I, [_P, _P, _I, _SZ, _I, _P, _P]),
    "avsep_pit_backward": (_I, [_I, _P, _P, _P, _P, _P, _I, _SZ, _I, _P, _P]),
    "avsep_bss_eval_workspace_bytes": (_SZ, [_I, _I, _I]),
    "avsep_bss_eval": (_I, [_P, C.c_longlong, C.c_longlong, _P, C.c_longlong, C.c_longlong, _I, _I, _I, _I, _P, _P, _P]),
    "avsep_step_create": (_I, [c_void_pp, _I, _I, _I, _I]),
    "avsep_step_destroy": (_I, [_P]),
    "avsep_step_train_host": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_step_train_host_async": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_step_train_dev": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_step_eval_host": (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _P, _P]),
    "avsep_step_eval_host_async": (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _P, _P]),
    "avsep_step_train_pcm16_async": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "avsep_pcm16_to_wave": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "avsep_step_eval_dev": (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _P, _P]),
    "avsep_step_buffer": (_P, [_P, _I]),
    "avsep_step_plan": (_P, [_P]),
    "avsep_step_set_mix_on_device": (_I, [_P, _I]),
    "avsep_step_set_fused_train": (_I, [_P, _I]),
    "avsep_step_set_graph": (_I, [_P, _I]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here = the library does not match the header
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.avsep_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise AvsepError(f"{what or 'avsep_b200'} failed (code {rc}): {last_error()}")


def ptr_array(ptrs):
    """A C array of void* from python ints (None -> NULL)."""
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
