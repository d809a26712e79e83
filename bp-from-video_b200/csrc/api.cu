// libbpv: error plumbing + version.
#include <stdarg.h>
#include <map>
#include <mutex>
#include <utility>
#include "common.cuh"

namespace bpv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, device): remember what each pair was raised
// to, so that repeated launches skip the driver call and a host process that drives several GPUs (or several
// threads) still configures every one of them.
int ensure_dyn_smem(const void* kernel, size_t bytes, bool max_carveout) {
  if (bytes <= 48 * 1024 && !max_carveout) return 0;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> raised;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = raised[{kernel, dev}];
  if (bytes <= have) return 0;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && max_carveout) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%zu B of dynamic shared memory): %s", bytes, cudaGetErrorString(e)); return (int)e; }
  have = bytes;
  return 0;
}
}  // namespace bpv

extern "C" int bpv_version(void) { return BPV_VERSION; }
extern "C" const char* bpv_last_error(void) { return bpv::g_err; }
extern "C" int bpv_sizeof_window_params(void) { return (int)sizeof(bpv_window_params); }

// L2 -> DRAM fetch granularity hint (32/64/128 B).  ROI rows are short, unaligned spans: a smaller
// granularity cuts the DRAM over-fetch around every row.  Device-wide limit of the primary context.
extern "C" int bpv_set_l2_fetch_granularity(int bytes) {
  cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
  if (e != cudaSuccess) { bpv::set_error("cudaDeviceSetLimit(L2 fetch granularity): %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}
extern "C" int bpv_get_l2_fetch_granularity(void) {
  size_t v = 0;
  if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return -1;
  return (int)v;
}


// ---- L2 discard of dead scratch (engine: proc_x / proc_y after F3 and F4 have consumed them) ----
// The processed windows are scratch between F2 and F3 / F4 (~80 MB per 8192-job step, rewritten every step).  Left dirty in
// L2 they are written back to DRAM while the next step's ROI sampling streams its frames through the cache: measured 64.5 us
// -> 69.7 us for the 8192-frame F1 launch with 80 MB of dirty lines in L2 (profiles/r2p_f1_dirty.txt).  discard.global.L2
// drops the lines without the write-back; their contents are undefined afterwards, which is exactly what scratch is.
namespace bpv {
__global__ void __launch_bounds__(256) l2_discard_kernel(uint8_t* base, long long lines) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < lines) asm volatile("discard.global.L2 [%0], 128;" :: "l"(base + i * 128) : "memory");
}
}  // namespace bpv

extern "C" int bpv_scratch_discard(void* ptr, int64_t bytes, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(ptr || bytes == 0, BPV_E_INVALID, "bpv_scratch_discard: NULL ptr");
  BPV_REQUIRE(bytes >= 0, BPV_E_INVALID, "bpv_scratch_discard: negative size");
  // only the 128-byte lines that lie entirely inside [ptr, ptr + bytes)
  const uintptr_t lo = ((uintptr_t)ptr + 127) & ~(uintptr_t)127, hi = ((uintptr_t)ptr + (uintptr_t)bytes) & ~(uintptr_t)127;
  if (hi <= lo) return 0;
  const long long lines = (long long)((hi - lo) >> 7);
  l2_discard_kernel<<<(unsigned)((lines + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint8_t*)lo, lines);
  return check_launch("bpv_scratch_discard");
}


// ---- FMA throughput probe (bench.py: measured FP64 / FP32 peaks for the roofline of the filter / spectrum families) ----
namespace bpv {
template <typename T>
__global__ void __launch_bounds__(256) probe_fma_kernel(long long iters, T* sink) {
  T a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
  const T m = (T)0.999999, c = (T)1e-7;
  for (long long it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
  }
  T sres = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) sres += a[i];
  if (!(sres == sres)) *sink = sres;      // never true: keeps the chain alive
}
}  // namespace bpv

extern "C" int64_t bpv_probe_fma(int32_t dtype, int64_t iters, int32_t blocks, void* sink, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(sink && iters > 0 && blocks > 0 && (dtype == 0 || dtype == 1), BPV_E_INVALID, "bpv_probe_fma: bad arguments");
  if (dtype == 1) probe_fma_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, (double*)sink);
  else probe_fma_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, (float*)sink);
  if (int rc = check_launch("bpv_probe_fma")) return -(int64_t)rc;
  return (int64_t)blocks * 256 * iters * 32;
}
