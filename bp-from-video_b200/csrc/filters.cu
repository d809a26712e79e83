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
  for (int k = 0; k < window; ++k) {
    const long long g = head - window + 1 + k;
    if (g < 0) continue;
    const double x = ring_t[g % cap];
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

// One CTA (64 threads) per filter: q/b vectors, Q = toeplitz + hankel (64x64, lda 65), Cholesky,
// two triangular solves, symmetric tap assembly — scipy.signal.firls (_fir_filter_design.py:1130-1171)
// with desired = [0,0,1,1,0,0], weight = 1.
constexpr int FIR_LDA = 65;
__device__ void firls_design_block(double fs, int taps, double min_freq, double max_freq, double df,
                                   double* __restrict__ out, double* smem) {
  const int M = (taps - 1) / 2, n = M + 1;  // n unknowns (<= 64)
  double* q = smem;                 // [128]
  double* rhs = q + 128;            // [64]
  double* A = rhs + 64;             // [64 * FIR_LDA]
  const int t = threadIdx.x;
  double fb[6];
  const bool ok = firls_bands(fs, min_freq, max_freq, df, fb);
  if (!ok) {
    for (int i = t; i < taps; i += blockDim.x) out[i] = nan_f64();
    return;
  }
  for (int i = t; i < taps; i += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < 3; ++b) acc += fb[2 * b + 1] * np_sinc(fb[2 * b + 1] * i) - fb[2 * b] * np_sinc(fb[2 * b] * i);
    q[i] = acc;
  }
  for (int i = t; i < n; i += blockDim.x) rhs[i] = fb[3] * np_sinc(fb[3] * i) - fb[2] * np_sinc(fb[2] * i);
  __syncthreads();
  for (int idx = t; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx % n;
    A[i * FIR_LDA + j] = q[i > j ? i - j : j - i] + q[i + j];
  }
  __syncthreads();
  // right-looking Cholesky, thread i owns row i (lower triangle)
  for (int k = 0; k < n; ++k) {
    const double dkk = sqrt(A[k * FIR_LDA + k]);
    __syncthreads();
    if (t == k) A[k * FIR_LDA + k] = dkk;
    if (t > k && t < n) A[t * FIR_LDA + k] /= dkk;
    __syncthreads();
    if (t > k && t < n) {
      const double lik = A[t * FIR_LDA + k];
      for (int j = k + 1; j <= t; ++j) A[t * FIR_LDA + j] -= lik * A[j * FIR_LDA + k];
    }
    __syncthreads();
  }
  // L z = rhs (column-oriented forward substitution)
  for (int k = 0; k < n; ++k) {
    if (t == k) rhs[k] /= A[k * FIR_LDA + k];
    __syncthreads();
    if (t > k && t < n) rhs[t] -= A[t * FIR_LDA + k] * rhs[k];
    __syncthreads();
  }
  // L^T a = z (backward)
  for (int k = n - 1; k >= 0; --k) {
    if (t == k) rhs[k] /= A[k * FIR_LDA + k];
    __syncthreads();
    if (t < k) rhs[t] -= A[k * FIR_LDA + t] * rhs[k];
    __syncthreads();
  }
  // coeffs = [a[M..1], 2 a0, a[1..M]]
  for (int i = t; i < taps; i += blockDim.x) {
    const int d = i < M ? M - i : i - M;
    out[i] = d == 0 ? 2.0 * rhs[0] : rhs[d];
  }
}

__global__ void __launch_bounds__(64) firls_from_fs_kernel(const double* __restrict__ fs, int taps, double min_freq,
                                                           double max_freq, double df, double* __restrict__ out) {
  extern __shared__ double smem[];
  firls_design_block(fs[blockIdx.x], taps, min_freq, max_freq, df, out + (long long)blockIdx.x * taps, smem);
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

__global__ void __launch_bounds__(64) job_firls_kernel(const double* __restrict__ ring_t, bpv_window_params p,
                                                       double* __restrict__ out) {
  extern __shared__ double smem[];
  __shared__ double fs_s;
  const int job = blockIdx.x;
  const int s = job / p.jobs_per_stream, j = job % p.jobs_per_stream;
  if (threadIdx.x == 0)
    fs_s = job_fs(ring_t + (long long)s * p.cap, p.cap, p.window, p.head0 + (long long)j * p.head_step, nullptr);
  __syncthreads();
  firls_design_block(fs_s, p.fir_taps, p.min_freq, p.max_freq, p.fir_df, out + (long long)job * p.fir_taps, smem);
}

constexpr size_t FIRLS_SMEM = (128 + 64 + 64 * FIR_LDA) * sizeof(double);

int launch_job_butter(const double* ring_t, const bpv_window_params& p, double* sos_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_butter_kernel<<<(J + 63) / 64, 64, 0, st>>>(ring_t, p, sos_out);
  return check_launch("job_butter_kernel");
}

int launch_job_firls(const double* ring_t, const bpv_window_params& p, double* taps_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_firls_kernel<<<J, 64, FIRLS_SMEM, st>>>(ring_t, p, taps_out);
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
  firls_from_fs_kernel<<<n, 64, FIRLS_SMEM, (cudaStream_t)stream>>>(fs, p->fir_taps, p->min_freq, p->max_freq, p->fir_df, taps_out);
  return check_launch("bpv_firls_design");
}
