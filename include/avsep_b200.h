/*
 * avsep_b200.h -- C-ABI of the B200-native mix-and-separate spectrogram hot path.
 *
 * The reference (abcqmars/audio-visual-sepatation-in-visual-agnostic-situtation) is pure
 * Python/PyTorch and has no FFI layer of its own; its "operator interface" for this path is
 * the set of Python call signatures listed in SURVEY.md section 8(b).  Each entry point below
 * names the reference function (file:line under the reference checkout) it replaces.  The Python
 * host side (avsep_b200/*.py) binds these with ctypes and re-exposes the reference's
 * signatures; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary
 *   - `dev` pointers are CUDA device pointers on the CURRENT device, fp32, contiguous,
 *     16-byte aligned; `host` pointers are host memory (pinned recommended)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call
 *     is asynchronous on that stream unless stated otherwise
 *   - return value: 0 on success, non-zero on failure; avsep_last_error() gives the message
 *     (thread-local).  There is NO CPU fallback: on a machine without an sm_100 GPU the
 *     compute entry points fail with AVSEP_ERR_NO_DEVICE.
 *   - borrowed inputs are never freed or written; outputs are caller-allocated
 */
#ifndef AVSEP_B200_H_
#define AVSEP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVSEP_ABI_VERSION 1

#define AVSEP_OK               0
#define AVSEP_ERR_INVALID      1   /* bad argument (shape / null pointer / unsupported size) */
#define AVSEP_ERR_CUDA         2   /* a CUDA runtime / driver call failed */
#define AVSEP_ERR_NO_DEVICE    3   /* no sm_100 device: the product path refuses to run     */
#define AVSEP_ERR_KERNEL       4   /* device-side watchdog tripped (pipeline deadlock guard) */

#define AVSEP_N_FFT   1022   /* arguments.py --stft_frame default; only value the kernels support */
#define AVSEP_HOP     256    /* arguments.py --stft_hop default */
#define AVSEP_N_BINS  512    /* n_fft/2 + 1 */
#define AVSEP_MAX_SRC 4      /* max sources N per mixture in the fused kernels */

/* loss kinds: models/criterion.py:28-49 */
#define AVSEP_LOSS_NONE 0
#define AVSEP_LOSS_L1   1
#define AVSEP_LOSS_L2   2
#define AVSEP_LOSS_BCE  3

/* activations: models/__init__.py:16-28 (only the two the path uses) */
#define AVSEP_ACT_NONE    0
#define AVSEP_ACT_SIGMOID 1

/* ---- library ------------------------------------------------------------------------- */
int         avsep_abi_version(void);
const char* avsep_last_error(void);
/* 0 if `device` (or the current device when device < 0) is compute capability 10.x */
int         avsep_device_check(int device);
/* number of kernels this library has launched in the calling process (for bench accounting) */
uint64_t    avsep_launch_count(void);

/* ---- STFT / iSTFT plan ----------------------------------------------------------------- */
/* Owns the split-precision DFT bases (bf16 hi/lo, window folded in), TMA descriptors and
 * the bf16 staging buffers for up to `max_signals` signals of `wave_len` samples. */
typedef struct avsep_plan avsep_plan;
int avsep_plan_create(avsep_plan** plan, int max_signals, int wave_len);
int avsep_plan_destroy(avsep_plan* plan);
/* frames T = 1 + wave_len / 256 (centre reflect padding), iSTFT length = 256 * (T - 1) */
int avsep_plan_frames(const avsep_plan* plan);
int avsep_plan_istft_len(const avsep_plan* plan);
size_t avsep_plan_workspace_bytes(const avsep_plan* plan);

/* K1. Replaces BaseDataset._stft / _mix_n_and_stft (dataset/base.py:142-147, 174-189) and
 * inference.get_spec (inference.py:221-226): librosa.stft(n_fft=1022, hop=256) -> |.|, angle.
 *   wave_dev [S, wave_len] -> mag_dev [S,512,T]; phase_dev / re_dev / im_dev [S,512,T] optional (NULL to skip).
 * Windowed DFT as a bf16x3 split-precision tcgen05 GEMM; magnitude / phase fused in the epilogue. */
int avsep_stft(avsep_plan* plan, const float* wave_dev, int S,
               float* mag_dev, float* phase_dev, float* re_dev, float* im_dev, void* stream);

/* Same transform, magnitudes of all S signals but (re, im) of the first reim_count signals only: re_dev / im_dev
 * [reim_count, 512, T].  This is the eval step's forward transform (signal-major waves: the first B signals are the
 * mixtures, whose complex spectrum the inverse transform needs; the sources only need |.| for the masks' targets). */
int avsep_stft_reim(avsep_plan* plan, const float* wave_dev, int S, float* mag_dev, float* re_dev, float* im_dev,
                    int reim_count, void* stream);

/* STFT of B mixtures given only their N sources (the hand-off of dataset/music.py:120-127, where the mixture IS the
 * sum of the sources): sources_dev [N, B, L] -> mag_dev [(N+1) B, 512, T] with the mixtures in planes [0, B) and source
 * n in planes [(1+n) B, (2+n) B); phase_dev (same layout) is optional.  The mixture waveform is never materialised:
 * with 2 <= N <= 4 and no phase output the mixture spectrum is taken as the sum of the source spectra (avsep_plan_set_mix_linear), otherwise the
 * staging kernel sums the sources on the fly.  Replaces BaseDataset._mix_n_and_stft (dataset/base.py:174-189) per batch. */
int avsep_stft_mix(avsep_plan* plan, const float* sources_dev, int B, int N, float* mag_dev, float* phase_dev, void* stream);
/* The evaluation form of avsep_stft_mix (main.py:213-249 needs the mixture's complex spectrum for the inverse transform):
 * magnitudes [(N+1) * B, 512, T] as above plus (re, im) of the B mixture spectra, re_dev / im_dev [B, 512, T].  With
 * 2 <= N <= 4 only the N source transforms run: mixture (re, im) = sum of the source spectra (dft_fold3m). */
int avsep_stft_mix_reim(avsep_plan* plan, const float* sources_dev, int B, int N, float* mag_dev, float* re_dev, float* im_dev,
                        void* stream);

/* K4. Replaces the reconstruction loop of calc_metrics (main.py:213-249) + utils.istft_reconstruction
 * (utils.py:101-104): un-warp each predicted mask 256->512 bins (grid_sample with warpgrid(warp=False)),
 * threshold (> mask_thres) if binary, multiply the mixture spectrum, inverse DFT (tcgen05 GEMM),
 * window, overlap-add, divide by the window envelope, trim n_fft/2, clip to [-1, 1].
 *   masks_dev  [N, B, 256, T]  predicted masks on the log-frequency grid (NULL with N = 0: mixture only)
 *   mag_dev    [B, 512, T], phase_dev [B, 512, T]   mixture spectrogram (phase_dev may be NULL when
 *              re_dev / im_dev [B,512,T] are given instead -- skips the sincos)
 *   wav_dev    [B, N+1, 256*(T-1)]  index 0 = mixture reconstruction, 1+n = source n
 *   log_freq = 0 means masks are already [N,B,512,T] (args.log_freq == 0 branch, main.py:221-222) */
int avsep_separate(avsep_plan* plan, const float* masks_dev, int N, int B,
                   const float* mag_dev, const float* phase_dev,
                   const float* re_dev, const float* im_dev,
                   int log_freq, int binary_mask, float mask_thres, int include_mix,
                   float* wav_dev, void* stream);

/* Plain utils.istft_reconstruction (utils.py:101-104) for S spectrograms, no mask:
 *   mag_dev, phase_dev [S,512,T] -> wav_dev [S, 256*(T-1)] */
int avsep_istft(avsep_plan* plan, const float* mag_dev, const float* phase_dev, int S,
                float* wav_dev, void* stream);

/* Debug/test hook: raw accumulator of the PLAIN forward (inverse = 0) or inverse (1) DFT GEMM for the operands
 * currently staged in the plan (forward: by avsep_stft after avsep_plan_set_fold(plan, 0)): out_dev [S, 1024, T] fp32,
 * row order = basis row order. */
int avsep_debug_raw_gemm(avsep_plan* plan, int inverse, int S, float* out_dev, void* stream);
/* The transform kernel alone with its product epilogue (forward: mag [S,512,T]; inverse: wav [S, 256*(T-1)]), so it can be
 * timed in isolation.  The folded forward kernel has no staged operand: it re-reads the waveform pointer of the last
 * avsep_stft call on this plan; the inverse runs on the operand the last avsep_separate staged. */
int avsep_debug_gemm_only(avsep_plan* plan, int inverse, int S, float* out_dev, void* stream);
/* copies the bf16 hi/lo basis (as fp32 sums) to host: out_host [1024, 1024] */
int avsep_debug_basis(avsep_plan* plan, int inverse, float* out_host);
/* staged forward operand (hi + lo) of signal s as fp32 on the host: out_host [T, 1024] */
int avsep_debug_staged_frames(avsep_plan* plan, int s, float* out_host);
/* enable = 1 (default): the forward transform runs as the folded kernels (real-input symmetry: half the tensor work;
 * the E / O operand is generated inside the kernel from the fp32 waveform, no staging kernel, no staged planes in HBM);
 * 0: the plain 1024 x 1024 DFT GEMM on a bf16 hi / lo staged waveform -- the in-library cross-check of the folded
 * kernels (also what avsep_debug_raw_gemm exercises). */
int avsep_plan_set_fold(avsep_plan* plan, int enable);
/* enable = 1 (default): the inverse transform of avsep_separate / avsep_istft (utils.py:101-104) runs as the FOLDED
 * GEMM (csrc/idft_fold.cuh): irfft output symmetry y[n] = P[n] + Q[n], y[1022 - n] = P[n] - Q[n] turns the
 * 1024 x 1024 contraction into two 512 x 512 ones (half the tensor work), overlap-add fused in the epilogue;
 * 0: the plain 1024 x 1024 inverse DFT GEMM (kept as the in-library cross-check of the folded kernel). */
int avsep_plan_set_inverse_fold(avsep_plan* plan, int enable);
/* enable = 1 (default): when a step forms the mixture on the device as the sum of its N = 2..4 sources
 * (avsep_step_set_mix_on_device; dataset/music.py:127 defines every mixture that way) and no phase output is asked for,
 * the mixture spectrum (magnitude, and (re, im) for the eval step) is taken as the sum of the source spectra -- the STFT
 * is linear -- so N transforms are computed instead of N + 1 (dft_fold3m).  0: the mixture waveform is summed and transformed like any other signal.
 * The two routes agree to fp32 rounding (tests/test_gpu_step.py). */
int avsep_plan_set_mix_linear(avsep_plan* plan, int enable);
/* SM partitioning: the persistent CTA-pair transforms of this plan (forward and inverse) launch at most sms / 2 pairs
 * (0 = the whole device, the default).  Used by pipelines that run the data-only front end (the STFT: it depends on the
 * waveforms, not on the model) of batch i + 1 on one partition of the SMs (a green context, or simply the SMs this
 * leaves free) while the HBM-bound synthesizer kernels of batch i run on the rest -- the overlap the reference gets
 * from its DataLoader workers (dataset/base.py:142-147 runs in them while main.py:104-139 trains).  Results do not
 * depend on the budget (tile order only). */
int avsep_plan_set_sm_budget(avsep_plan* plan, int sms);
/* enable = 1 (default): avsep_separate builds the inverse-transform operand (un-warp, threshold, mask x mixture spectrum;
 * main.py:213-246) with the shared-memory staged kernel (cp.async loads, TMA tensor stores; csrc/spec_prep.cuh) whenever
 * (re, im) of the mixture are given and T % 4 == 0; 0: always the register-tile kernel (the in-library cross-check: the
 * two produce identical operands). */
int avsep_plan_set_spec_staged(avsep_plan* plan, int enable);

/* ---- K2: log-frequency warp + derivations ------------------------------------------------- */
/* Replaces utils.warpgrid + F.grid_sample (utils.py:12-26; main.py:61-66, 216-220): separable
 * 2-tap bilinear resample of the frequency axis (and the mild time-axis resample grid_sample
 * applies with align_corners=False).  x_dev [S, Hin, T] -> y_dev [S, Hout, T].
 *   warp = 1: linear->log (Hin 512 -> Hout 256); warp = 0: log->linear (256 -> 512). */
int avsep_warp(const float* x_dev, int S, int Hin, int Hout, int T, int warp, float* y_dev, void* stream);

/* Host-side tap tables of that resample (no GPU needed; used by the CPU tests):
 *   kind 0 = warp rows, 1 = un-warp rows, 2 = linear (columns); i0_host / frac_host [n_out] */
int avsep_warp_taps_host(int kind, int n_in, int n_out, int* i0_host, float* frac_host);

/* Replaces NetWrapper.prepare (main.py:51-95), tensor part:
 *   mag_mix_dev [B,512,T]; mags_dev[n] -> [B,512,T] for n < N (separate tensors, as batch_data['mags'])
 *   outputs (all [B,256,T], any may be NULL): mag_mix_w, mags_w[n], log_mag_mix, weights, gt_masks[n] */
int avsep_prepare(const float* mag_mix_dev, const float* const* mags_dev, int N, int B, int T,
                  int log_freq, int weighted_loss, int binary_mask,
                  float* mag_mix_w_dev, float* const* mags_w_dev, float* log_mag_mix_dev,
                  float* weights_dev, float* const* gt_masks_dev, void* stream);

/* ---- K3: synthesizer mask (+ sigmoid, + loss) and backward ------------------------------- */
/* Replaces InnerProd.forward / Bias.forward (models/synthesizer_net.py:12-19, 46-51) for all N
 * sources of a mixture in ONE pass over feat_sound, the caller's activate(.., 'sigmoid')
 * (models/__init__.py:16-18) and, optionally, BaseLoss.forward list form (models/criterion.py:10-49).
 *   feat_dev  [B,K,HW]   vecs_dev [B,N,K]   scale_dev [K] or NULL (Bias)   bias_dev [1]
 *   out_dev   [N,B,HW]   activation(z)
 *   loss (loss_kind != NONE): gt_dev [N,B,HW], weight_dev [B,HW] or NULL (= ones),
 *        loss_dev [1] receives mean_n mean_{b,hw} w * l(pred_n, gt_n)
 *   scratch_dev: >= avsep_synth_scratch_bytes(B, N, K, HW) bytes of device scratch */
size_t avsep_synth_scratch_bytes(int B, int N, int K, int HW);
int avsep_synth_forward(const float* feat_dev, const float* vecs_dev, const float* scale_dev,
                        const float* bias_dev, int B, int N, int K, int HW, int activation,
                        float* out_dev, int loss_kind, const float* gt_dev, const float* weight_dev,
                        float* loss_dev, void* scratch_dev, void* stream);

/* Backward of the above.  Either (fused) the loss gradient is recomputed from the saved
 * activations:  dz = gscale * dL/dpred * act'(z)   with pred_dev/gt_dev/weight_dev and the upstream
 * scalar gradient gscale_dev [1];  or (given) dz_dev [N,B,HW] is supplied (plain InnerProd autograd).
 * Outputs (any may be NULL): dfeat_dev [B,K,HW], dvecs_dev [B,N,K], dscale_dev [K], dbias_dev [1].
 * torch semantics are kept: BCE gradient uses max(p(1-p), 1e-12) and is exactly 0 where p saturates. */
int avsep_synth_backward(const float* feat_dev, const float* vecs_dev, const float* scale_dev,
                         int B, int N, int K, int HW, int activation, int loss_kind,
                         const float* pred_dev, const float* gt_dev, const float* weight_dev,
                         const float* gscale_dev, const float* dz_dev,
                         float* dfeat_dev, float* dvecs_dev, float* dscale_dev, float* dbias_dev,
                         void* scratch_dev, void* stream);

/* Fused training step of the synthesizer: forward + activation + list-form loss + full backward with ONE read of
 * feat_dev from HBM (K <= 32; larger K runs the two kernels above).  Gradients are those of
 * gscale_dev[0] * loss (NULL = 1, i.e. loss.backward()).  pred_dev [N,B,HW] receives the activations. */
int avsep_synth_train(const float* feat_dev, const float* vecs_dev, const float* scale_dev, const float* bias_dev,
                      int B, int N, int K, int HW, int activation, int loss_kind,
                      const float* gt_dev, const float* weight_dev, const float* gscale_dev,
                      float* pred_dev, float* loss_dev, float* dfeat_dev, float* dvecs_dev, float* dscale_dev,
                      float* dbias_dev, void* scratch_dev, void* stream);

/* InnerProd.forward_nosum (synthesizer_net.py:21-26): out [B,K,HW] = vec*scale (x) feat + bias */
int avsep_synth_nosum(const float* feat_dev, const float* vec_dev, const float* scale_dev,
                      const float* bias_dev, int B, int K, int HW, float* out_dev, void* stream);
/* InnerProd.forward_pixelwise (synthesizer_net.py:29-38): feats_img [B,K,P] x feat [B,K,HW] -> [B,P,HW] */
int avsep_synth_pixelwise(const float* feat_dev, const float* feats_img_dev, const float* scale_dev,
                          const float* bias_dev, int B, int K, int P, int HW, float* out_dev, void* stream);

/* ---- criterion as a stand-alone op (models/criterion.py:6-49) ----------------------------- */
/* preds_dev[n], targets_dev[n] -> [numel]; weight_dev [numel] or NULL; loss_dev [1].
 * scratch: >= avsep_loss_scratch_bytes(N, numel). */
size_t avsep_loss_scratch_bytes(int N, size_t numel);
int avsep_loss_forward(int loss_kind, const float* const* preds_dev, const float* const* targets_dev,
                       const float* weight_dev, int N, size_t numel, float* loss_dev,
                       void* scratch_dev, void* stream);
int avsep_loss_backward(int loss_kind, const float* const* preds_dev, const float* const* targets_dev,
                        const float* weight_dev, int N, size_t numel, const float* gscale_dev,
                        float* const* dpreds_dev, void* stream);

/* x_dev[0..n) *= gscale_dev[0], skipped on the device when the scalar is exactly 1 (no host sync, no traffic) */
int avsep_scale_if_needed(float* x_dev, size_t n, const float* gscale_dev, void* stream);

/* PitWrapper(F.binary_cross_entropy) (models/criterion.py:74-231): per-sample S x S loss matrix.
 *   preds_dev, targets_dev, weights_dev [B, P, S] (sources innermost, as the reference stacks them)
 *   lossmat_dev [B, S, S]: lossmat[b][i][j] = mean_p w[b,p,i] * l(pred[b,p,j], tgt[b,p,i])
 * The S! permutation search runs on the host side of the binding (S <= 4). */
int avsep_pit_lossmat(int loss_kind, const float* preds_dev, const float* targets_dev,
                      const float* weights_dev, int B, size_t P, int S, float* lossmat_dev, void* stream);

/* SoP++ audio-only step, forward (SoP++/main.py:227-236 + PitWrapper._opt_perm_loss, models/criterion.py:138-176): all N
 * synthesizer masks of every mixture AND the per-sample N x N PIT loss matrix
 *   lossmat[b][i][j] = mean_HW weight * l(act(z_j), gt_i)
 * in ONE pass over feat_sound.  Layouts as avsep_synth_forward (out / gt [N,B,HW], weight [B,HW] or NULL, H*W % 4 == 0);
 * lossmat_dev [B,N,N] feeds avsep_pit_search.  scratch_dev: avsep_synth_scratch_bytes(B, N, K, HW). */
int avsep_synth_forward_pit(const float* feat_dev, const float* vecs_dev, const float* scale_dev, const float* bias_dev,
                            int B, int N, int K, int HW, int activation, float* out_dev, int loss_kind,
                            const float* gt_dev, const float* weight_dev, float* lossmat_dev, void* scratch_dev,
                            void* stream);
/* y[j][b][:] = x[sel[b][j]][b][:] * (rowscale ? rowscale[b] * factor : 1) for x, y [N,B,P]; sel_dev [B,N] int32 or NULL
 * (identity).  Re-orders source planes per sample by a PIT permutation (PitWrapper.reorder_tensor in the [N,B,...] layout)
 * and folds per-sample upstream gradients into the loss weights for the backward of the audio-only step. */
int avsep_plane_gather(const float* x_dev, const int* sel_dev, const float* rowscale_dev, float factor, int N, int B,
                       size_t P, float* y_dev, void* stream);
/* PitWrapper._fast_pit (models/criterion.py:113-136) on the device: for every sample the S! permutations in
 * itertools.permutations order, cost = mean_i lossmat[b][i][perm[i]] in fp32, first strict minimum wins.
 *   lossmat_dev [B,S,S] -> loss_dev [B], perm_dev [B,S] (int32), inv_perm_dev [B,S] (int32; what avsep_pit_backward takes).
 * No D2H copy, no host synchronisation. */
int avsep_pit_search(const float* lossmat_dev, int B, int S, float* loss_dev, int* perm_dev, int* inv_perm_dev, void* stream);
/* PitWrapper.reorder_tensor (models/criterion.py:180-200): y[b, ..., s] = x[b, ..., perm[b][s]] for x [B, P, S]. */
int avsep_pit_reorder(const float* x_dev, const int* perm_dev, int B, size_t P, int S, float* y_dev, void* stream);
/* Backward of the PIT loss  loss[b] = (1/S) sum_i lossmat[b][i][perm_b[i]]  w.r.t. preds:
 *   inv_perm_dev [B, S] int32: inv_perm[b][j] = the target index matched with prediction j;  gloss_dev [B]. */
int avsep_pit_backward(int loss_kind, const float* preds_dev, const float* targets_dev, const float* weights_dev,
                       const int* inv_perm_dev, const float* gloss_dev, int B, size_t P, int S, float* dpreds_dev,
                       void* stream);

/* ---- end-to-end host-buffer entry points (what bench.py's `e2e` times) --------------------- */
/* ---- separation metrics ---------------------------------------------------------------- */
/* Replaces the per-clip metric call of calc_metrics (main.py:252-275): asteroid.metrics.get_metrics(mix, clean, estimate,
 * sample_rate=11025, metrics_list=['sdr','sir','sar','si_sdr']) = mir_eval.separation.bss_eval_sources(compute_permutation=
 * False) (BSS-eval v3, 512-tap distortion filters) + pb_bss_eval si_sdr, for all B clips at once, in fp64, on the device:
 * direct lag-0..511 cross-correlations, block-Toeplitz Levinson solves of the (N*512)^2 Gram systems, projection energies as
 * quadratic forms (csrc/bss_eval.cu).  No FFT, no dense factorisation, no library call.
 *   ref_dev / est_dev : reference sources / estimates, float32 (is_double = 0) or float64 (1); element (b, n, t) sits at
 *                       base[b * stride_b + n * stride_n + t], t < L  (strides in elements: waveforms of a step and its
 *                       wav output are scored in place, without gathering)
 *   out_dev           : double [4][B][N] = sdr, sir, sar, si_sdr per clip and source (the reference averages over sources)
 *   workspace_dev     : avsep_bss_eval_workspace_bytes(B, N, L) bytes */
size_t avsep_bss_eval_workspace_bytes(int B, int N, int L);
int avsep_bss_eval(const void* ref_dev, long long ref_stride_b, long long ref_stride_n,
                   const void* est_dev, long long est_stride_b, long long est_stride_n, int is_double,
                   int B, int N, int L, double* out_dev, void* workspace_dev, void* stream);

/* Config C2 train step through host buffers: H2D of the waveforms, STFT of all (N+1) signals,
 * prepare, synth + sigmoid + weighted BCE, full backward, D2H of the loss.
 *   waves_host [N+1, B, wave_len], SIGNAL-major: plane 0 = the B mixtures, plane 1+n = source n of every mixture
 *              (pinned memory recommended); [N, B, wave_len] = the sources only after
 *              avsep_step_set_mix_on_device(step, 1).  The same layout holds for waves_dev / waves_host of every
 *              step entry point below (train_host, train_host_async, train_dev, eval_host, eval_dev).  A
 *              mixture-major [B, N+1, wave_len] buffer has the same element count and would be accepted silently
 *              with wrong mixture / source assignment: callers must transpose.
 *   vecs_host  [B, N, K]
 *   feat_dev   [B, K, 256*T] stays on the device (it is the U-Net's output in the reference and
 *              never exists on the host there); dfeat_dev [B,K,256*T] receives its gradient
 *   loss_host  [1].  Synchronises `stream` before returning. */
typedef struct avsep_step avsep_step;
int avsep_step_create(avsep_step** step, int B, int N, int K, int wave_len);
int avsep_step_destroy(avsep_step* step);
int avsep_step_train_host(avsep_step* step, const float* waves_host, const float* vecs_host,
                          const float* feat_dev, const float* scale_dev, const float* bias_dev,
                          float* dfeat_dev, float* dvecs_host, float* loss_host, void* stream);
/* Same, but only enqueued on `stream` (no synchronisation): with two steps on two streams the H2D copy of
 * step i+1 overlaps the kernels of step i.  loss_host / dvecs_host must be pinned and are valid once the
 * stream has been synchronised. */
int avsep_step_train_host_async(avsep_step* step, const float* waves_host, const float* vecs_host,
                                const float* feat_dev, const float* scale_dev, const float* bias_dev,
                                float* dfeat_dev, float* dvecs_host, float* loss_host, void* stream);
/* Opt-in 16-bit PCM hand-off for the same step: the host passes the clips as the int16 samples the wav files hold,
 * pcm_host [N+1, B, wave_len] (or [N, B, wave_len] with mix_on_device), plus optional per-clip gains gains_host
 * [N+1, B] (NULL = 1.0; the train-time amplitude augmentation of dataset/base.py:165-167).  The device applies the
 * loader's float32 operations in the loader's order -- x = int16 / 32768; x *= gain; clip to [-1, 1]; x /= N
 * (dataset/base.py:156-172, dataset/music.py:120) -- so the waveform the STFT sees is bit-identical to the float32 one
 * the reference's loader produces from the same samples, at half the H2D bytes.  Enqueue only, like
 * avsep_step_train_host_async. */
int avsep_step_train_pcm16_async(avsep_step* step, const int16_t* pcm_host, const float* gains_host,
                                 const float* vecs_host, const float* feat_dev, const float* scale_dev,
                                 const float* bias_dev, float* dfeat_dev, float* dvecs_host, float* loss_host,
                                 void* stream);
/* the conversion alone on device buffers: pcm_dev [clips, L] int16 (+ gains_dev [clips] or NULL) -> wave_dev [clips, L] */
int avsep_pcm16_to_wave(const int16_t* pcm_dev, const float* gains_dev, int clips, int L, int num_mix,
                        float* wave_dev, void* stream);
/* same step with device-resident inputs, asynchronous (what bench.py's `value` times) */
int avsep_step_train_dev(avsep_step* step, const float* waves_dev, const float* vecs_dev,
                         const float* feat_dev, const float* scale_dev, const float* bias_dev,
                         float* dfeat_dev, float* dvecs_dev, float* dscale_dev, float* dbias_dev,
                         float* loss_dev, void* stream);
/* Config C1/C3 eval step: STFT -> prepare -> synth + sigmoid -> un-warp/threshold/mask -> iSTFT.
 *   wav_out [B, N+1, 256*(T-1)] on the host (eval_host) or the device (eval_dev). */
int avsep_step_eval_host(avsep_step* step, const float* waves_host, const float* vecs_host,
                         const float* feat_dev, const float* scale_dev, const float* bias_dev,
                         int binary_mask, float mask_thres, float* wav_host, void* stream);
/* eval_host without the final synchronisation (two steps on two streams overlap copies and kernels); wav_host may be
 * NULL: the separated waveforms then stay on the device (avsep_step_buffer id 13) for a device-side evaluator. */
int avsep_step_eval_host_async(avsep_step* step, const float* waves_host, const float* vecs_host,
                               const float* feat_dev, const float* scale_dev, const float* bias_dev,
                               int binary_mask, float mask_thres, float* wav_host, void* stream);
int avsep_step_eval_dev(avsep_step* step, const float* waves_dev, const float* vecs_dev,
                        const float* feat_dev, const float* scale_dev, const float* bias_dev,
                        int binary_mask, float mask_thres, float* wav_dev, void* stream);
/* device views of the step's intermediates (valid until the next call): 0 mag [(N+1)*B,512,T] (signal-major like
 * the waveforms), 1 mix re, 2 mix im [B,512,T], 3 mag_mix_w, 4 log_mag_mix, 5 weights [B,256,T], 6 gt [N,B,256,T],
 * 7 pred masks [N,B,256,T], 8 loss [1], 9 dvecs [B,N,K], 10 dscale [K], 11 dbias [1], 12 warped source magnitudes
 * [N,B,256,T], 13 wav [B,N+1,256*(T-1)] (eval steps); any other id returns NULL */
const float* avsep_step_buffer(const avsep_step* step, int which);
/* enable = 1: the `waves` argument of every step call holds only the N sources, [N, B, wave_len]; the mixture
 * (dataset/music.py:127: the sum of the sources) is formed on the GPU inside the staging kernel, which saves
 * 1/(N+1) of the H2D bytes.  Default 0: waves = [N+1, B, wave_len] with the mixture at index 0. */
int avsep_step_set_mix_on_device(avsep_step* step, int enable);
/* enable = 1: avsep_step_train_dev / avsep_step_eval_dev (and the host variants built on them) capture their kernel sequence into a CUDA
 * graph on the second call with an unchanged argument tuple and replay it afterwards (one cudaGraphLaunch per
 * step).  Only on explicitly created streams (the legacy default stream cannot be captured).  Default 0. */
int avsep_step_set_graph(avsep_step* step, int enable);
/* enable = 0: run the separate synthesizer forward and backward kernels instead of the fused training kernel
 * (A/B timing and tests); default 1. */
int avsep_step_set_fused_train(avsep_step* step, int enable);
/* the STFT / iSTFT plan the step owns (borrowed) */
avsep_plan* avsep_step_plan(avsep_step* step);

#ifdef __cplusplus
}
#endif
#endif  /* AVSEP_B200_H_ */
