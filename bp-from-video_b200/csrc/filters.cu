// Filter-design kernels: one Butterworth SOS / one least-squares FIR per window job (the R ROIs of a
// job share the job's timestamps, hence fs, hence the filter).  float64 throughout.
#include "filters.cuh"

namespace bpv {

// ---- job sampling rate: Signal.get_fs (signal_data.py:55-58) over the job's window --------------
// fs = 1 / nanmean(diff(x[finite])) == (m-1) / (x_last - x_first).  After an INTERP_* step the
// reference uses 1/linspace-step, the same quantity up to rounding (signal_processor.py:211,218).
__device__ inline double job_fs(const double* __restrict__ ring_t, int cap, int window, long long head, int* m_out) {
  double first = 0, last = 0;
  int m = 0;
  long long g = head - window + 1;
  int k = 0;
  if (g < 0) { k = (int)(-g < window ? -g : window); g = 0; }     // samples before the stream started read as NaN
  int slot = (int)(g % cap);
  for (; k < window; ++k) {
    const double x = ring_t[slot];
    if (++slot == cap) slot = 0;
    if (isfinite(x)) { if (m == 0) first = x; last = x; ++m; }
  }
  if (m_out) *m_out = m;
  return m >= 2 ? 1.0 / ((last - first) / (double)(m - 1)) : nan_f64();
}

__global__ void butter_from_fs_kernel(const double* __restrict__ fs, int n, int order, double min_freq, double max_freq,
                                      double min_bw, double* __restrict__ sos_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sos[MAX_SOS * 6];
  butter_bandpass_sos(fs[i], order, min_freq, max_freq, min_bw, sos);
  for (int k = 0; k < order * 6; ++k) sos_out[(long long)i * order * 6 + k] = sos[k];
}

// One WARP per filter — scipy.signal.firls (_fir_filter_design.py:1130-1171) with desired = [0,0,1,1,0,0],
// weight = 1.
//
// scipy solves Q a = b with Q = toeplitz(q[:M+1]) + hankel(q[:M+1], q[M:]) (Cholesky) and then mirrors a into
// the taps h = [a[M..1], 2 a0, a[1..M]].  Written for h directly, that is the symmetric positive-definite
// TOEPLITZ system  T h = y,  T[u][v] = q[|u-v|] (numtaps x numtaps), y = [b[M..1], b0, b[1..M]]:
// row i >= 0 of T h is 2 q[i] a0 + sum_{j>=1} (q[|i-j|] + q[i+j]) a[j] = (Q a)[i].  T is better conditioned
// than Q (cond 10 at 30 fps, 7e4 at 8.7 fps) and a Toeplitz solve is O(n^2): Levinson-Durbin, 126 steps of
// two dot products + two AXPYs of growing length, instead of a 64^3/3 factorisation.  Agreement with scipy's
// firls: <= 8e-13 relative over fs 8.7..240 Hz (checked against numpy in tests/test_window_gpu.py).
// Lane l owns elements l, l+32, l+64, l+96 of the forward vector f and of the solution x (registers); f is
// mirrored in shared memory (ping-pong) because every step needs it reversed.
constexpr int FIRLS_WARPS = 4;                      // filters per CTA
constexpr int FIRLS_THREADS = 32 * FIRLS_WARPS;
constexpr int FIRLS_WS = 4 * 128;                   // doubles of shared memory per warp: r | y | f ping | f pong
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ void firls_design_warp(double fs, int taps, double min_freq, double max_freq, double df,
                                  double* __restrict__ out, double* __restrict__ zi_out, double* smem) {
  const int M = (taps - 1) / 2, n = taps;
  double* r = smem;                 // [128] first column of T: q[0 .. taps-1]
  double* y = r + 128;              // [128] symmetric right-hand side
  double* fbuf = y + 128;           // [2][128] forward vector, ping-pong
  const int lane = threadIdx.x & 31;
  double fb[6];
  const bool ok = firls_bands(fs, min_freq, max_freq, df, fb);
  if (!ok) {
    for (int i = lane; i < taps; i += 32) out[i] = nan_f64();
    return;
  }
  for (int i = lane; i < 128; i += 32) {
    double acc = 0.0, bv = 0.0;
    if (i < taps) {
      for (int b = 0; b < 3; ++b) acc += fb[2 * b + 1] * np_sinc(fb[2 * b + 1] * i) - fb[2 * b] * np_sinc(fb[2 * b] * i);
      const int d = i < M ? M - i : i - M;        // y[i] = b[|i - M|], b[d] = f3 sinc(f3 d) - f2 sinc(f2 d)
      bv = fb[3] * np_sinc(fb[3] * d) - fb[2] * np_sinc(fb[2] * d);
    }
    r[i] = acc;
    y[i] = bv;
    fbuf[i] = 0.0; fbuf[128 + i] = 0.0;
  }
  __syncwarp();
  double f[4] = {0, 0, 0, 0}, x[4] = {0, 0, 0, 0};
  const double r0inv = 1.0 / r[0];
  if (lane == 0) { f[0] = r0inv; x[0] = y[0] * r0inv; fbuf[0] = r0inv; }
  __syncwarp();
  for (int k = 1; k < n; ++k) {
    const double* fo = fbuf + ((k - 1) & 1) * 128;   // f of step k-1 (length k)
    double* fn = fbuf + (k & 1) * 128;               // f of step k (length k+1)
    const int mmax = k >> 5;                         // element groups that can be non-empty (warp uniform)
    double ef = 0.0, ex = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m > mmax) break;
      const int i = lane + 32 * m;
      if (i < k) { const double rk = r[k - i]; ef = fma(rk, f[m], ef); ex = fma(rk, x[m], ex); }
    }
    ef = warp_sum_d(ef); ex = warp_sum_d(ex);
    const double dinv = 1.0 / (1.0 - ef * ef);
    const double g = y[k] - ex;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m > mmax) break;
      const int i = lane + 32 * m;
      if (i <= k) {
        const double frev = i >= 1 ? fo[k - i] : 0.0;          // bb[i] = f_old[k - i], bb[0] = 0
        f[m] = (f[m] - ef * frev) * dinv;                       // fb[k] = 0 is already in the register
        fn[i] = f[m];
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m > mmax) break;
      const int i = lane + 32 * m;
      if (i <= k) x[m] = fma(g, fn[k - i], x[m]);               // x += (y[k] - ex) * reversed(f_new)
    }
    __syncwarp();
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = lane + 32 * m;
    if (i < n) out[i] = x[m];
  }
  if (zi_out) {
    // lfilter_zi(b, [1]) = suffix sums of b[1:] (scipy/signal/_signaltools.py:4440-4463): zi[i] = sum_{k>i} b[k].
    // Lane l holds taps l, l+32, l+64, l+96; suffix sums per 32-chunk by a shuffle scan, chunks chained.
    double carry = 0.0;                                   // sum of all taps in higher chunks
#pragma unroll
    for (int m = 3; m >= 0; --m) {
      const int i = lane + 32 * m;
      const double v = i < n ? x[m] : 0.0;
      double incl = v;                                    // inclusive suffix sum over lanes >= lane
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
      }
      if (i < n - 1) zi_out[i] = (incl - v) + carry;      // strictly-greater taps
      carry += __shfl_sync(0xffffffffu, incl, 0);
    }
  }
}

__global__ void __launch_bounds__(FIRLS_THREADS) firls_from_fs_kernel(const double* __restrict__ fs, int n, int taps, double min_freq,
                                                                      double max_freq, double df, double* __restrict__ out) {
  extern __shared__ double smem[];
  const int w = threadIdx.x >> 5, job = blockIdx.x * FIRLS_WARPS + w;
  if (job >= n) return;
  firls_design_warp(fs[job], taps, min_freq, max_freq, df, out + (long long)job * taps, nullptr, smem + w * FIRLS_WS);
}

// ---- per-job design from the ring timestamps (used by the window pipeline) ----------------------
__global__ void job_butter_kernel(const double* __restrict__ ring_t, bpv_window_params p, double* __restrict__ sos_out) {
  const int job = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = p.S * p.jobs_per_stream;
  if (job >= J) return;
  const int s = job / p.jobs_per_stream, j = job % p.jobs_per_stream;
  const double fs = job_fs(ring_t + (long long)s * p.cap, p.cap, p.window, p.head0 + (long long)j * p.head_step, nullptr);
  double sos[MAX_SOS * 6];
  if (isfinite(fs)) butter_bandpass_sos(fs, p.butter_order, p.min_freq, p.max_freq, p.butter_min_bw, sos);
  else for (int k = 0; k < p.butter_order * 6; ++k) sos[k] = nan_f64();
  for (int k = 0; k < p.butter_order * 6; ++k) sos_out[(long long)job * p.butter_order * 6 + k] = sos[k];
}

__global__ void __launch_bounds__(FIRLS_THREADS) job_firls_kernel(const double* __restrict__ ring_t, bpv_window_params p,
                                                                  double* __restrict__ out) {
  extern __shared__ double smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x * FIRLS_WARPS + w;
  if (job >= p.S * p.jobs_per_stream) return;
  const int s = job / p.jobs_per_stream, j = job % p.jobs_per_stream;
  // fs of the job's window, warp cooperative (first / last finite timestamp and their count)
  const double* rt = ring_t + (long long)s * p.cap;
  const long long g0 = p.head0 + (long long)j * p.head_step - p.window + 1;
  int lo = 0x7fffffff, hi = -1, cnt = 0;
  const int kmin = g0 < 0 ? (int)(-g0 < p.window ? -g0 : p.window) : 0;
  const int slot0 = (int)(((g0 % p.cap) + p.cap) % p.cap);
  for (int k = lane; k < p.window; k += 32) {
    int slot = slot0 + k; if (slot >= p.cap) slot -= p.cap;
    if (k >= kmin && isfinite(rt[slot])) { lo = lo < k ? lo : k; hi = k; ++cnt; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo < l2 ? lo : l2; hi = hi > h2 ? hi : h2; cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  double fs = nan_f64();
  if (cnt >= 2) fs = 1.0 / ((rt[(slot0 + hi) % p.cap] - rt[(slot0 + lo) % p.cap]) / (double)(cnt - 1));
  firls_design_warp(fs, p.fir_taps, p.min_freq, p.max_freq, p.fir_df, out + (long long)job * 256, out + (long long)job * 256 + 128,
                    smem + w * FIRLS_WS);
}

constexpr size_t FIRLS_SMEM = (size_t)FIRLS_WARPS * FIRLS_WS * sizeof(double);

int launch_job_butter(const double* ring_t, const bpv_window_params& p, double* sos_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_butter_kernel<<<(J + 63) / 64, 64, 0, st>>>(ring_t, p, sos_out);
  return check_launch("job_butter_kernel");
}

int launch_job_firls(const double* ring_t, const bpv_window_params& p, double* taps_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_firls_kernel<<<(J + FIRLS_WARPS - 1) / FIRLS_WARPS, FIRLS_THREADS, FIRLS_SMEM, st>>>(ring_t, p, taps_out);
  return check_launch("job_firls_kernel");
}

int check_filter_params(const bpv_window_params* p, const char* who) {
  BPV_REQUIRE(p, BPV_E_INVALID, "%s: NULL params", who);
  BPV_REQUIRE(p->butter_order >= 1 && p->butter_order <= MAX_SOS, BPV_E_TOO_LARGE, "%s: butter_order must be 1..16", who);
  BPV_REQUIRE(p->fir_taps >= 3 && p->fir_taps <= MAX_TAPS && (p->fir_taps & 1), BPV_E_TOO_LARGE,
              "%s: fir_taps must be odd and <= 127", who);
  return 0;
}

}  // namespace bpv

extern "C" int bpv_butter_sos_design(const double* fs, int32_t n, const bpv_window_params* p, double* sos_out, void* stream) {
  using namespace bpv;
  if (int rc = check_filter_params(p, "bpv_butter_sos_design")) return rc;
  BPV_REQUIRE(fs && sos_out && n >= 0, BPV_E_INVALID, "bpv_butter_sos_design: bad arguments");
  if (n == 0) return 0;
  butter_from_fs_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(fs, n, p->butter_order, p->min_freq, p->max_freq,
                                                                        p->butter_min_bw, sos_out);
  return check_launch("bpv_butter_sos_design");
}

extern "C" int bpv_firls_design(const double* fs, int32_t n, const bpv_window_params* p, double* taps_out, void* stream) {
  using namespace bpv;
  if (int rc = check_filter_params(p, "bpv_firls_design")) return rc;
  BPV_REQUIRE(fs && taps_out && n >= 0, BPV_E_INVALID, "bpv_firls_design: bad arguments");
  if (n == 0) return 0;
  firls_from_fs_kernel<<<(n + FIRLS_WARPS - 1) / FIRLS_WARPS, FIRLS_THREADS, FIRLS_SMEM, (cudaStream_t)stream>>>(fs, n, p->fir_taps, p->min_freq, p->max_freq, p->fir_df, taps_out);
  return check_launch("bpv_firls_design");
}
