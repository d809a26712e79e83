// Device ring buffers of raw samples (sg_raw): the batched twin of Signal.add_sample /
// SignalGroup.add_samples (signal_data.py:31-35, 94-98).  deque(maxlen) semantics are a ring
// addressed by the global sample index; the NaN prefill (signal_data.py:18-19) is the caller's
// initial fill plus the "negative global index reads NaN" rule of the window kernels.
#include "common.cuh"

namespace bpv {
__global__ void ring_push_kernel(double* __restrict__ ring_t, double* __restrict__ ring_y, int S, int R, int cap,
                                 long long g0, int T, const double* __restrict__ ts, const double* __restrict__ values) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over S*T*(R+1)
  const long long total = (long long)S * T * (R + 1);
  if (i >= total) return;
  const int c = (int)(i % (R + 1));
  const long long st = i / (R + 1);
  const int t = (int)(st % T);
  const long long s = st / T;
  const int slot = (int)((g0 + t) % cap);
  if (c == R) ring_t[s * cap + slot] = ts[s * T + t];
  else ring_y[(s * R + c) * cap + slot] = values[(s * T + t) * R + c];
}
}  // namespace bpv

extern "C" int bpv_ring_push(double* ring_t, double* ring_y, int32_t S, int32_t R, int32_t cap,
                             int64_t g0, int32_t T, const double* ts, const double* values, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(ring_t && ring_y && ts && values, BPV_E_INVALID, "bpv_ring_push: NULL pointer");
  BPV_REQUIRE(S > 0 && R > 0 && cap > 0 && T > 0 && T <= cap && g0 >= 0, BPV_E_INVALID, "bpv_ring_push: bad sizes");
  const long long total = (long long)S * T * (R + 1);
  ring_push_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ring_t, ring_y, S, R, cap, g0, T, ts, values);
  return check_launch("bpv_ring_push");
}
