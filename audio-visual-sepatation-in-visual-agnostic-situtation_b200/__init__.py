"""avsep_b200 -- B200-native (sm_100a) mix-and-separate spectrogram hot path.

Drop-in for the hot path of abcqmars/audio-visual-sepatation-in-visual-agnostic-situtation:
batched STFT (tcgen05 split-precision DFT GEMM), log-frequency warp + prepare derivations,
synthesizer mask + sigmoid + weighted loss (forward / backward), un-warp + mask + iSTFT
overlap-add.  Hand-written CUDA behind a C-ABI (include/avsep_b200.h); this package is the
ctypes binding plus host-side mirrors of the reference's Python signatures.

The directory name carries the upstream repo name (hyphens), so import it through the
`avsep_b200` alias module at the repository root:  `import avsep_b200`.
"""
from . import _lib
from ._lib import AvsepError, LIB_PATH
from . import ops
from .ops import (Plan, clear_plans, device_check, get_plan, istft, launch_count, list_loss, pcm16_to_wave, pit_loss, pit_lossmat, pit_reorder, pit_search, separate, stft, stft_mix,
                  synth_masks, synth_masks_loss, synth_masks_pit_loss, synth_nosum, synth_pixelwise)
from . import modules
from .modules import BCELoss, BaseLoss, Bias, InnerProd, L1Loss, L2Loss, PitWrapper, activate
from . import hotpath
from .hotpath import (Step, StepPipeline, _mix_n_and_stft, _stft, assemble_windows, calc_metrics, clip_windows, get_spec, istft_reconstruction, prepare,
                      clear_inferdata_cache, prepare_inferdata, recover_aud, recover_aud_arrays,
                      separate_long_clip, stft_batch, warp, warpgrid)
from . import parallel
from . import bss_metrics
from .bss_metrics import bss_eval_batch, bss_eval_batch_torch

__all__ = [
    "AvsepError", "LIB_PATH", "StepPipeline", "clear_inferdata_cache", "ops", "modules", "hotpath", "parallel", "bss_metrics", "bss_eval_batch", "bss_eval_batch_torch",
    "Plan", "get_plan", "clear_plans", "device_check", "launch_count",
    "stft", "stft_mix", "istft", "pcm16_to_wave", "separate", "warp", "warpgrid", "prepare", "calc_metrics", "istft_reconstruction",
    "_stft", "_mix_n_and_stft", "stft_batch", "get_spec", "prepare_inferdata", "recover_aud", "recover_aud_arrays", "Step", "clip_windows", "assemble_windows", "separate_long_clip",
    "synth_masks", "synth_masks_loss", "synth_masks_pit_loss", "synth_nosum", "synth_pixelwise", "list_loss", "pit_loss", "pit_lossmat",
    "InnerProd", "Bias", "BaseLoss", "BCELoss", "L1Loss", "L2Loss", "PitWrapper", "activate",
]
