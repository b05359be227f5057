// End-to-end steps of the hot path behind one C call (what bench.py times):
//   train (config C2): STFT of the (N+1) signals -> prepare -> synth + sigmoid + weighted BCE -> backward
//   eval  (config C1/C3): STFT -> prepare -> synth + sigmoid (+ loss) -> un-warp / threshold / mask -> iSTFT
// Host-buffer variants add the H2D copy of the waveforms / K-vectors and the D2H copy of the result
// inside the call.  feat_sound is the U-Net's output in the reference and never exists on the host
// there, so it stays a device pointer.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

struct avsep_plan;

namespace avsep {
int prepare_impl(const float* mag_mix, const float* const* mags, int N, int B, int T, int log_freq, int weighted,
                 int binary, float* mix_w, float* const* mags_w, float* log_mix, float* weights, float* const* gt,
                 cudaStream_t stream);
int synth_forward_impl(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int N, int K, int HW,
                       int act, float* out, int loss_kind, const float* gt, const float* weight, float* loss, void* scratch,
                       cudaStream_t stream);
int stft_impl(avsep_plan* pl, const float* wave_dev, int S, float* mag_dev, float* phase_dev, float* re_dev, float* im_dev,
              int reim_count, void* stream, int mix_B, int mix_N);
int plan_generation(const avsep_plan* pl);
bool synth_train_supported(int K, int HW, const float* dfeat);
int synth_train_impl(const float* feat, const float* vecs, const float* scale, const float* bias, int B, int N, int K, int HW,
                     int act, int loss_kind, const float* gt, const float* weight, const float* gscale, float* pred, float* loss,
                     float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch, cudaStream_t stream);
int synth_backward_impl(const float* feat, const float* vecs, const float* scale, int B, int N, int K, int HW, int act,
                        int loss_kind, const float* pred, const float* gt, const float* weight, const float* gscale,
                        const float* dz, float* dfeat, float* dvecs, float* dscale, float* dbias, void* scratch,
                        cudaStream_t stream);
}  // namespace avsep

// One captured CUDA graph of a step's kernel sequence, keyed on the argument tuple it was captured with
struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  const void* key[16] = {};
  int launches = 0;            // kernels inside the graph (for the launch counter)
  void reset() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    memset(key, 0, sizeof(key));
  }
};

struct avsep_step {
  int B = 0, N = 0, K = 0, L = 0, T = 0, HW = 0, out_len = 0;
  int mix_on_device = 0;      // waves arguments hold only the N sources; the mixture is summed on the GPU
  int no_fused_train = 0;     // debugging / A-B timing: run the separate forward and backward kernels
  int use_graph = 0;          // replay the train / eval step as a CUDA graph (non-default streams only)
  GraphSlot train_graph, eval_graph;
  avsep_plan* plan = nullptr;
  float* waves = nullptr;     // [N+1, B, L]  (index 0 = mixture)
  float* vecs = nullptr;      // [B, N, K]
  float* mag = nullptr;       // [(N+1) B, 512, T]
  float* mix_re = nullptr;    // [B, 512, T]
  float* mix_im = nullptr;
  float* mix_w = nullptr;     // [B, 256, T]
  float* log_mix = nullptr;
  float* weights = nullptr;
  float* mags_w = nullptr;    // [N, B, 256, T]
  float* gt = nullptr;        // [N, B, 256, T]
  float* pred = nullptr;      // [N, B, 256, T]
  float* loss = nullptr;      // [1]
  float* one = nullptr;       // [1] = 1.0f (upstream gradient of err.backward())
  float* dvecs = nullptr;     // [B, N, K]
  float* dscale = nullptr;    // [K]
  float* dbias = nullptr;     // [1]
  float* wav = nullptr;       // [B, N+1, out_len]
  int16_t* pcm = nullptr;     // [N+1, B, L] 16-bit PCM staging of the opt-in hand-off (allocated on first use)
  float* gains = nullptr;     // [N+1, B] per-clip amplitude gains of that hand-off
  void* scratch = nullptr;
};

using namespace avsep;

static int step_forward_common(avsep_step* st, const float* waves_dev, const float* vecs_dev, const float* feat_dev,
                               const float* scale_dev, const float* bias_dev, bool need_reim, cudaStream_t stream,
                               bool with_synth = true) {
  const int B = st->B, N = st->N;
  int rc = stft_impl(st->plan, waves_dev, (N + 1) * B, st->mag, nullptr, need_reim ? st->mix_re : nullptr,
                     need_reim ? st->mix_im : nullptr, B, stream, st->mix_on_device ? B : 0, st->mix_on_device ? N : 0);
  if (rc) return rc;
  const size_t sig = (size_t)AVSEP_N_BINS * st->T;
  const size_t img = (size_t)st->HW;
  const float* mags[AVSEP_MAX_SRC];
  float* mags_w[AVSEP_MAX_SRC];
  float* gts[AVSEP_MAX_SRC];
  for (int n = 0; n < N; ++n) {
    mags[n] = st->mag + (size_t)(1 + n) * B * sig;
    mags_w[n] = st->mags_w + (size_t)n * B * img;
    gts[n] = st->gt + (size_t)n * B * img;
  }
  rc = prepare_impl(st->mag, mags, N, B, st->T, 1, 1, 1, st->mix_w, mags_w, st->log_mix, st->weights, gts, stream);
  if (rc || !with_synth) return rc;
  return synth_forward_impl(feat_dev, vecs_dev, scale_dev, bias_dev, B, N, st->K, st->HW, AVSEP_ACT_SIGMOID, st->pred,
                            AVSEP_LOSS_BCE, st->gt, st->weights, st->loss, st->scratch, stream);
}

// 16-bit PCM hand-off (dataset/base.py:150-172 + dataset/music.py:120): the loader's float waveform is
//   audio = int16 / 32768 (wav decode) ; audio *= gain (train-time augmentation, base.py:165-167) ; clip to [-1, 1] ; / num_mix
// -- the same fp32 operations in the same order here, so the device waveform is bit-identical to the one the reference's
// loader would have produced from the same PCM samples, at half the H2D bytes.
__global__ void __launch_bounds__(256) pcm16_to_wave_kernel(const int16_t* __restrict__ pcm, const float* __restrict__ gains,
                                                            float* __restrict__ wave, int L, float num_mix) {
  const int clip = blockIdx.y;
  const float g = gains ? __ldg(gains + clip) : 1.0f;
  const int16_t* src = pcm + (size_t)clip * L;
  float* dst = wave + (size_t)clip * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    float x = (float)src[i] * (1.0f / 32768.0f);       // exact
    x = x * g;                                         // numpy float32 multiply (round to nearest)
    x = fminf(fmaxf(x, -1.0f), 1.0f);
    dst[i] = __fdiv_rn(x, num_mix);                    // numpy float32 / int -> float32 division
  }
}

// H2D copies of the PCM samples (and gains), enqueued back to back with the caller's other input copies; the conversion
// kernel is launched by pcm16_convert AFTER all copies of the step (a kernel between two copies of a stream keeps the
// next step's copy from overlapping this step's kernels)
static int pcm16_copy(avsep_step* st, const int16_t* pcm_host, const float* gains_host, cudaStream_t stream) {
  const int signals = st->N + (st->mix_on_device ? 0 : 1);
  const size_t n = (size_t)signals * st->B * st->L;
  if (!st->pcm) {
    AVSEP_CUDA(cudaMalloc(&st->pcm, sizeof(int16_t) * (size_t)(st->N + 1) * st->B * st->L));
    AVSEP_CUDA(cudaMalloc(&st->gains, sizeof(float) * (size_t)(st->N + 1) * st->B));
  }
  AVSEP_CUDA(cudaMemcpyAsync(st->pcm, pcm_host, sizeof(int16_t) * n, cudaMemcpyHostToDevice, stream));
  if (gains_host)
    AVSEP_CUDA(cudaMemcpyAsync(st->gains, gains_host, sizeof(float) * (size_t)signals * st->B, cudaMemcpyHostToDevice, stream));
  return AVSEP_OK;
}

static int pcm16_convert(avsep_step* st, bool with_gains, cudaStream_t stream) {
  const int signals = st->N + (st->mix_on_device ? 0 : 1);
  dim3 grid(std::min(ceil_div(st->L, 256 * 8), 64), signals * st->B);
  pcm16_to_wave_kernel<<<grid, 256, 0, stream>>>(st->pcm, with_gains ? st->gains : nullptr, st->waves, st->L, (float)st->N);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

// A step is a fixed kernel sequence.  With graphs enabled the sequence is captured once per argument tuple (on its second
// call, after the lazily-created tap tables / operand buffers exist) and replayed with one cudaGraphLaunch, which removes
// the per-kernel launch gaps (~7 % of a 0.65 ms train step, ~4 % of the 1.2 ms eval step).
template <typename Body>
static int run_graphed(avsep_step* st, GraphSlot& g, const void* const (&key)[16], cudaStream_t stream, Body body) {
  if (!st->use_graph || stream == nullptr || stream == cudaStreamLegacy || stream == cudaStreamPerThread) return body();
  if (g.exec && memcmp(key, g.key, sizeof(g.key)) == 0) {
    AVSEP_CUDA(cudaGraphLaunch(g.exec, stream));
    count_launch(g.launches);
    return AVSEP_OK;
  }
  if (memcmp(key, g.key, sizeof(g.key)) != 0) {            // first call with these arguments: run directly (warm-up)
    g.reset();
    memcpy(g.key, key, sizeof(g.key));
    return body();
  }
  // second call with the same arguments: capture, instantiate, launch
  const uint64_t before = avsep_launch_count();
  AVSEP_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  int rc = body();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(stream, &graph);
  if (rc || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    st->use_graph = 0;                                     // capture is not possible here: fall back to plain launches
    if (rc) return rc;
    return body();
  }
  ce = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    g.exec = nullptr;
    st->use_graph = 0;
    cudaGetLastError();
    return body();
  }
  g.launches = (int)(avsep_launch_count() - before);
  AVSEP_CUDA(cudaGraphLaunch(g.exec, stream));
  return AVSEP_OK;
}

extern "C" {

int avsep_step_destroy(avsep_step* st) {
  if (!st) return AVSEP_OK;
  st->train_graph.reset();
  st->eval_graph.reset();
  avsep_plan_destroy(st->plan);
  float* bufs[] = {st->waves, st->vecs, st->mag, st->mix_re, st->mix_im, st->mix_w, st->log_mix, st->weights, st->mags_w,
                   st->gt, st->pred, st->loss, st->one, st->dvecs, st->dscale, st->dbias, st->wav};
  for (float* b : bufs) cudaFree(b);
  cudaFree(st->scratch);
  cudaFree(st->pcm);
  cudaFree(st->gains);
  delete st;
  return AVSEP_OK;
}

int avsep_step_create(avsep_step** step, int B, int N, int K, int wave_len) {
  AVSEP_REQUIRE(step != nullptr, "avsep_step_create: step is null");
  *step = nullptr;
  AVSEP_REQUIRE(B >= 1 && N >= 1 && N <= AVSEP_MAX_SRC && K >= 1 && K <= 256, "avsep_step_create: bad B=%d N=%d K=%d", B, N, K);
  int rc = require_sm100();
  if (rc) return rc;
  avsep_step* st = new avsep_step();
  st->B = B; st->N = N; st->K = K; st->L = wave_len;
  rc = avsep_plan_create(&st->plan, (N + 1) * B, wave_len);
  if (rc) { delete st; return rc; }
  st->T = avsep_plan_frames(st->plan);
  st->HW = 256 * st->T;
  st->out_len = avsep_plan_istft_len(st->plan);
  const size_t sig = (size_t)AVSEP_N_BINS * st->T, img = (size_t)st->HW;
  struct { float** p; size_t n; } allocs[] = {
      {&st->waves, (size_t)(N + 1) * B * wave_len}, {&st->vecs, (size_t)B * N * K}, {&st->mag, (size_t)(N + 1) * B * sig},
      {&st->mix_re, (size_t)B * sig}, {&st->mix_im, (size_t)B * sig}, {&st->mix_w, (size_t)B * img},
      {&st->log_mix, (size_t)B * img}, {&st->weights, (size_t)B * img}, {&st->mags_w, (size_t)N * B * img},
      {&st->gt, (size_t)N * B * img}, {&st->pred, (size_t)N * B * img}, {&st->loss, 1}, {&st->one, 1},
      {&st->dvecs, (size_t)B * N * K}, {&st->dscale, (size_t)K}, {&st->dbias, 1},
      {&st->wav, (size_t)B * (N + 1) * st->out_len}};
  for (auto& a : allocs) {
    cudaError_t e = cudaMalloc(a.p, a.n * sizeof(float));
    if (e != cudaSuccess) {
      avsep_step_destroy(st);
      return set_error(AVSEP_ERR_CUDA, "avsep_step_create: cudaMalloc of %zu bytes failed: %s", a.n * sizeof(float), cudaGetErrorString(e));
    }
  }
  cudaError_t e = cudaMalloc(&st->scratch, avsep_synth_scratch_bytes(B, N, K, st->HW));
  const float one = 1.0f;
  if (e == cudaSuccess) e = cudaMemcpy(st->one, &one, sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    avsep_step_destroy(st);
    return set_error(AVSEP_ERR_CUDA, "avsep_step_create: %s", cudaGetErrorString(e));
  }
  *step = st;
  return AVSEP_OK;
}

static int train_dev_body(avsep_step* st, const float* waves_dev, const float* vecs_dev, const float* feat_dev,
                          const float* scale_dev, const float* bias_dev, float* dfeat_dev, float* dvecs_dev, float* dscale_dev,
                          float* dbias_dev, float* loss_dev, cudaStream_t stream) {
  const bool fused = dfeat_dev != nullptr && synth_train_supported(st->K, st->HW, dfeat_dev) && !st->no_fused_train;
  int rc = step_forward_common(st, waves_dev, vecs_dev, feat_dev, scale_dev, bias_dev, false, stream, !fused);
  if (rc) return rc;
  if (fused)
    rc = synth_train_impl(feat_dev, vecs_dev, scale_dev, bias_dev, st->B, st->N, st->K, st->HW, AVSEP_ACT_SIGMOID, AVSEP_LOSS_BCE,
                          st->gt, st->weights, nullptr, st->pred, st->loss, dfeat_dev, dvecs_dev ? dvecs_dev : st->dvecs,
                          dscale_dev ? dscale_dev : st->dscale, dbias_dev ? dbias_dev : st->dbias, st->scratch, stream);
  else
    rc = synth_backward_impl(feat_dev, vecs_dev, scale_dev, st->B, st->N, st->K, st->HW, AVSEP_ACT_SIGMOID, AVSEP_LOSS_BCE,
                           st->pred, st->gt, st->weights, st->one, nullptr, dfeat_dev, dvecs_dev ? dvecs_dev : st->dvecs,
                           dscale_dev ? dscale_dev : st->dscale, dbias_dev ? dbias_dev : st->dbias, st->scratch, stream);
  if (rc) return rc;
  if (loss_dev) AVSEP_CUDA(cudaMemcpyAsync(loss_dev, st->loss, sizeof(float), cudaMemcpyDeviceToDevice, stream));
  return AVSEP_OK;
}

int avsep_step_train_dev(avsep_step* st, const float* waves_dev, const float* vecs_dev, const float* feat_dev,
                         const float* scale_dev, const float* bias_dev, float* dfeat_dev, float* dvecs_dev, float* dscale_dev,
                         float* dbias_dev, float* loss_dev, void* stream_) {
  AVSEP_REQUIRE(st && waves_dev && vecs_dev && feat_dev, "avsep_step_train_dev: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  // the key also carries the plan's variant generation: avsep_plan_set_* after a capture must not replay the old kernels
  const void* key[16] = {waves_dev, vecs_dev, feat_dev, scale_dev, bias_dev, dfeat_dev, dvecs_dev, dscale_dev, dbias_dev, loss_dev,
                         (const void*)(uintptr_t)(st->mix_on_device * 2 + st->no_fused_train), (const void*)stream,
                         (const void*)(uintptr_t)plan_generation(st->plan)};
  return run_graphed(st, st->train_graph, key, stream, [&]() {
    return train_dev_body(st, waves_dev, vecs_dev, feat_dev, scale_dev, bias_dev, dfeat_dev, dvecs_dev, dscale_dev, dbias_dev, loss_dev, stream);
  });
}

int avsep_step_set_graph(avsep_step* st, int enable) {
  AVSEP_REQUIRE(st != nullptr, "avsep_step_set_graph: step is null");
  st->use_graph = enable ? 1 : 0;
  if (!enable) {
    st->train_graph.reset();
    st->eval_graph.reset();
  }
  return AVSEP_OK;
}

int avsep_step_train_host_async(avsep_step* st, const float* waves_host, const float* vecs_host, const float* feat_dev,
                                const float* scale_dev, const float* bias_dev, float* dfeat_dev, float* dvecs_host,
                                float* loss_host, void* stream_) {
  AVSEP_REQUIRE(st && waves_host && vecs_host && feat_dev && loss_host, "avsep_step_train_host_async: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  AVSEP_CUDA(cudaMemcpyAsync(st->waves, waves_host, sizeof(float) * (size_t)(st->N + (st->mix_on_device ? 0 : 1)) * st->B * st->L, cudaMemcpyHostToDevice, stream));
  AVSEP_CUDA(cudaMemcpyAsync(st->vecs, vecs_host, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyHostToDevice, stream));
  int rc = avsep_step_train_dev(st, st->waves, st->vecs, feat_dev, scale_dev, bias_dev, dfeat_dev, nullptr, nullptr, nullptr,
                                nullptr, stream);
  if (rc) return rc;
  AVSEP_CUDA(cudaMemcpyAsync(loss_host, st->loss, sizeof(float), cudaMemcpyDeviceToHost, stream));
  if (dvecs_host)
    AVSEP_CUDA(cudaMemcpyAsync(dvecs_host, st->dvecs, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyDeviceToHost, stream));
  return AVSEP_OK;
}

int avsep_step_train_pcm16_async(avsep_step* st, const int16_t* pcm_host, const float* gains_host, const float* vecs_host,
                                 const float* feat_dev, const float* scale_dev, const float* bias_dev, float* dfeat_dev,
                                 float* dvecs_host, float* loss_host, void* stream_) {
  AVSEP_REQUIRE(st && pcm_host && vecs_host && feat_dev && loss_host, "avsep_step_train_pcm16_async: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = pcm16_copy(st, pcm_host, gains_host, stream);
  if (rc) return rc;
  AVSEP_CUDA(cudaMemcpyAsync(st->vecs, vecs_host, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyHostToDevice, stream));
  rc = pcm16_convert(st, gains_host != nullptr, stream);
  if (rc) return rc;
  rc = avsep_step_train_dev(st, st->waves, st->vecs, feat_dev, scale_dev, bias_dev, dfeat_dev, nullptr, nullptr, nullptr,
                            nullptr, stream);
  if (rc) return rc;
  AVSEP_CUDA(cudaMemcpyAsync(loss_host, st->loss, sizeof(float), cudaMemcpyDeviceToHost, stream));
  if (dvecs_host)
    AVSEP_CUDA(cudaMemcpyAsync(dvecs_host, st->dvecs, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyDeviceToHost, stream));
  return AVSEP_OK;
}

int avsep_pcm16_to_wave(const int16_t* pcm_dev, const float* gains_dev, int clips, int L, int num_mix, float* wave_dev, void* stream) {
  AVSEP_REQUIRE(pcm_dev && wave_dev && clips >= 0 && clips <= 65535 && L >= 1 && num_mix >= 1, "avsep_pcm16_to_wave: bad argument");
  int rc = require_sm100();
  if (rc) return rc;
  if (clips == 0) return AVSEP_OK;
  dim3 grid(std::min(ceil_div(L, 256 * 8), 64), clips);
  pcm16_to_wave_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pcm_dev, gains_dev, wave_dev, L, (float)num_mix);
  AVSEP_LAUNCH_CHECK();
  return AVSEP_OK;
}

int avsep_step_train_host(avsep_step* st, const float* waves_host, const float* vecs_host, const float* feat_dev,
                          const float* scale_dev, const float* bias_dev, float* dfeat_dev, float* dvecs_host, float* loss_host,
                          void* stream) {
  int rc = avsep_step_train_host_async(st, waves_host, vecs_host, feat_dev, scale_dev, bias_dev, dfeat_dev, dvecs_host, loss_host, stream);
  if (rc) return rc;
  AVSEP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return AVSEP_OK;
}

int avsep_step_eval_dev(avsep_step* st, const float* waves_dev, const float* vecs_dev, const float* feat_dev,
                        const float* scale_dev, const float* bias_dev, int binary_mask, float mask_thres, float* wav_dev,
                        void* stream_) {
  AVSEP_REQUIRE(st && waves_dev && vecs_dev && feat_dev, "avsep_step_eval_dev: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint32_t thres_bits;
  memcpy(&thres_bits, &mask_thres, sizeof(thres_bits));
  const void* key[16] = {waves_dev, vecs_dev, feat_dev, scale_dev, bias_dev, wav_dev, (const void*)(uintptr_t)thres_bits,
                         (const void*)(uintptr_t)(st->mix_on_device * 2 + (binary_mask ? 1 : 0)), (const void*)stream,
                         (const void*)(uintptr_t)plan_generation(st->plan)};
  return run_graphed(st, st->eval_graph, key, stream, [&]() {
    int rc = step_forward_common(st, waves_dev, vecs_dev, feat_dev, scale_dev, bias_dev, true, stream);
    if (rc) return rc;
    return avsep_separate(st->plan, st->pred, st->N, st->B, nullptr, nullptr, st->mix_re, st->mix_im, 1, binary_mask, mask_thres, 1,
                          wav_dev ? wav_dev : st->wav, stream);
  });
}

int avsep_step_eval_host_async(avsep_step* st, const float* waves_host, const float* vecs_host, const float* feat_dev,
                               const float* scale_dev, const float* bias_dev, int binary_mask, float mask_thres, float* wav_host,
                               void* stream_) {
  AVSEP_REQUIRE(st && waves_host && vecs_host && feat_dev, "avsep_step_eval_host_async: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  AVSEP_CUDA(cudaMemcpyAsync(st->waves, waves_host, sizeof(float) * (size_t)(st->N + (st->mix_on_device ? 0 : 1)) * st->B * st->L, cudaMemcpyHostToDevice, stream));
  AVSEP_CUDA(cudaMemcpyAsync(st->vecs, vecs_host, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyHostToDevice, stream));
  int rc = avsep_step_eval_dev(st, st->waves, st->vecs, feat_dev, scale_dev, bias_dev, binary_mask, mask_thres, st->wav, stream);
  if (rc) return rc;
  if (wav_host)      // NULL: the separated waveforms stay on the device (avsep_step_buffer 13) for a device-side evaluator
    AVSEP_CUDA(cudaMemcpyAsync(wav_host, st->wav, sizeof(float) * (size_t)st->B * (st->N + 1) * st->out_len, cudaMemcpyDeviceToHost, stream));
  return AVSEP_OK;
}

int avsep_step_eval_host(avsep_step* st, const float* waves_host, const float* vecs_host, const float* feat_dev,
                         const float* scale_dev, const float* bias_dev, int binary_mask, float mask_thres, float* wav_host,
                         void* stream_) {
  AVSEP_REQUIRE(st && waves_host && vecs_host && feat_dev && wav_host, "avsep_step_eval_host: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  AVSEP_CUDA(cudaMemcpyAsync(st->waves, waves_host, sizeof(float) * (size_t)(st->N + (st->mix_on_device ? 0 : 1)) * st->B * st->L, cudaMemcpyHostToDevice, stream));
  AVSEP_CUDA(cudaMemcpyAsync(st->vecs, vecs_host, sizeof(float) * (size_t)st->B * st->N * st->K, cudaMemcpyHostToDevice, stream));
  int rc = avsep_step_eval_dev(st, st->waves, st->vecs, feat_dev, scale_dev, bias_dev, binary_mask, mask_thres, st->wav, stream);
  if (rc) return rc;
  AVSEP_CUDA(cudaMemcpyAsync(wav_host, st->wav, sizeof(float) * (size_t)st->B * (st->N + 1) * st->out_len, cudaMemcpyDeviceToHost, stream));
  AVSEP_CUDA(cudaStreamSynchronize(stream));
  return AVSEP_OK;
}

avsep_plan* avsep_step_plan(avsep_step* st) { return st ? st->plan : nullptr; }

int avsep_step_set_fused_train(avsep_step* st, int enable) {
  AVSEP_REQUIRE(st != nullptr, "avsep_step_set_fused_train: step is null");
  st->no_fused_train = enable ? 0 : 1;
  return AVSEP_OK;
}

int avsep_step_set_mix_on_device(avsep_step* st, int enable) {
  AVSEP_REQUIRE(st != nullptr, "avsep_step_set_mix_on_device: step is null");
  st->mix_on_device = enable ? 1 : 0;
  return AVSEP_OK;
}

const float* avsep_step_buffer(const avsep_step* st, int which) {
  if (!st) return nullptr;
  switch (which) {
    case 0: return st->mag;
    case 1: return st->mix_re;
    case 2: return st->mix_im;
    case 3: return st->mix_w;
    case 4: return st->log_mix;
    case 5: return st->weights;
    case 6: return st->gt;
    case 7: return st->pred;
    case 8: return st->loss;
    case 9: return st->dvecs;
    case 10: return st->dscale;
    case 11: return st->dbias;
    case 12: return st->mags_w;
    case 13: return st->wav;
    default: return nullptr;
  }
}

}  // extern "C"
