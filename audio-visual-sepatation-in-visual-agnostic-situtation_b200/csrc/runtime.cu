// Library runtime: error reporting, device gating (sm_100 only, no CPU fallback), launch
// accounting and the device cache of the separable warp tap tables.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "warp_taps.h"

namespace avsep {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::mutex g_dev_mutex;
static std::map<int, std::pair<int, int>> g_dev_info;  // device -> (major, sm count)

static int device_info(int device, int* major, int* sms) {
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  auto it = g_dev_info.find(device);
  if (it == g_dev_info.end()) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(AVSEP_ERR_NO_DEVICE, "no usable CUDA device (%s); avsep_b200 has no CPU fallback",
                       cudaGetErrorString(e));
    }
    it = g_dev_info.emplace(device, std::make_pair(prop.major, prop.multiProcessorCount)).first;
  }
  *major = it->second.first;
  *sms = it->second.second;
  return AVSEP_OK;
}

int require_sm100() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(AVSEP_ERR_NO_DEVICE, "no CUDA device (%s); avsep_b200 has no CPU fallback",
                     cudaGetErrorString(e));
  }
  int major = 0, sms = 0;
  int rc = device_info(dev, &major, &sms);
  if (rc) return rc;
  if (major != 10)
    return set_error(AVSEP_ERR_NO_DEVICE,
                     "device %d is compute capability %d.x; avsep_b200 kernels are built for sm_100a only", dev, major);
  return AVSEP_OK;
}

int sm_count() {
  int dev = 0, major = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (device_info(dev, &major, &sms)) return 148;
  return sms;
}

// ---------------------------------------------------------------------------------------------
// tap tables
// ---------------------------------------------------------------------------------------------
void compute_taps(int kind, int n_in, int n_out, int* i0, float* frac) {
  // np.linspace(-1, 1, n): arange(n) * step + start with the last element forced to stop
  const double step = (n_out > 1) ? 2.0 / (double)(n_out - 1) : 0.0;
  for (int o = 0; o < n_out; ++o) {
    double y = (o == n_out - 1 && n_out > 1) ? 1.0 : -1.0 + (double)o * step;
    double gd;
    if (kind == TAP_WARP_Y) gd = (std::pow(21.0, (y + 1.0) / 2.0) - 11.0) / 10.0;
    else if (kind == TAP_UNWARP_Y) gd = std::log(y * 10.0 + 11.0) / std::log(21.0) * 2.0 - 1.0;
    else gd = y;
    const float g = (float)gd;                                   // grid.astype(np.float32)
    const float i = ((g + 1.0f) * (float)n_in - 1.0f) / 2.0f;    // ATen grid_sampler_unnormalize
    const float fl = std::floor(i);
    i0[o] = (int)fl;
    frac[o] = i - fl;
  }
}

struct TapEntry {
  int* i0 = nullptr;
  float* frac = nullptr;
};
static std::mutex g_tap_mutex;
static std::map<std::tuple<int, int, int, int>, TapEntry> g_taps;

int get_taps(int kind, int n_in, int n_out, Taps* out) {
  int dev = 0;
  AVSEP_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tap_mutex);
  auto key = std::make_tuple(dev, kind, n_in, n_out);
  auto it = g_taps.find(key);
  if (it == g_taps.end()) {
    std::vector<int> hi(n_out);
    std::vector<float> hf(n_out);
    compute_taps(kind, n_in, n_out, hi.data(), hf.data());
    TapEntry e;
    AVSEP_CUDA(cudaMalloc(&e.i0, sizeof(int) * n_out));
    AVSEP_CUDA(cudaMalloc(&e.frac, sizeof(float) * n_out));
    // one-time, synchronous: the tables are immutable afterwards and shared by all streams
    AVSEP_CUDA(cudaMemcpy(e.i0, hi.data(), sizeof(int) * n_out, cudaMemcpyHostToDevice));
    AVSEP_CUDA(cudaMemcpy(e.frac, hf.data(), sizeof(float) * n_out, cudaMemcpyHostToDevice));
    it = g_taps.emplace(key, e).first;
  }
  out->i0 = it->second.i0;
  out->frac = it->second.frac;
  return AVSEP_OK;
}

}  // namespace avsep

extern "C" {

int avsep_abi_version(void) { return AVSEP_ABI_VERSION; }
const char* avsep_last_error(void) { return avsep::last_error().c_str(); }
uint64_t avsep_launch_count(void) { return avsep::g_launches.load(); }

int avsep_device_check(int device) {
  if (device < 0) return avsep::require_sm100();
  int major = 0, sms = 0;
  int rc = avsep::device_info(device, &major, &sms);
  if (rc) return rc;
  if (major != 10)
    return avsep::set_error(AVSEP_ERR_NO_DEVICE, "device %d is compute capability %d.x, need 10.x", device, major);
  return AVSEP_OK;
}

// CPU-testable export of the tap tables (tests compare it with the oracle / the reference warpgrid)
int avsep_warp_taps_host(int kind, int n_in, int n_out, int* i0_host, float* frac_host) {
  if (kind < 0 || kind > 2 || n_in < 1 || n_out < 1 || !i0_host || !frac_host)
    return avsep::set_error(AVSEP_ERR_INVALID, "avsep_warp_taps_host: bad arguments");
  avsep::compute_taps(kind, n_in, n_out, i0_host, frac_host);
  return AVSEP_OK;
}

}  // extern "C"
