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

// Block-cooperative version (one window per CTA): threads stride the window, shared atomics pick the first /
// last finite timestamp and count them.  s_i = 3 ints of shared memory.
__device__ inline double job_fs_block(const double* __restrict__ ring_t, int cap, int window, long long head, int* s_i) {
  if (threadIdx.x == 0) { s_i[0] = 0x7fffffff; s_i[1] = -1; s_i[2] = 0; }
  __syncthreads();
  const long long g0 = head - window + 1;
  int lo = 0x7fffffff, hi = -1, cnt = 0;
  for (int k = threadIdx.x; k < window; k += blockDim.x) {
    const long long g = g0 + k;
    if (g < 0) continue;
    if (isfinite(ring_t[g % cap])) { lo = lo < k ? lo : k; hi = k; ++cnt; }
  }
  if (cnt) { atomicMin(&s_i[0], lo); atomicMax(&s_i[1], hi); atomicAdd(&s_i[2], cnt); }
  __syncthreads();
  const int m = s_i[2];
  if (m < 2) return nan_f64();
  const double first = ring_t[(g0 + s_i[0]) % cap], last = ring_t[(g0 + s_i[1]) % cap];
  return 1.0 / ((last - first) / (double)(m - 1));
}

__global__ void butter_from_fs_kernel(const double* __restrict__ fs, int n, int order, double min_freq, double max_freq,
                                      double min_bw, double* __restrict__ sos_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sos[MAX_SOS * 6];
  butter_bandpass_sos(fs[i], order, min_freq, max_freq, min_bw, sos);
  for (int k = 0; k < order * 6; ++k) sos_out[(long long)i * order * 6 + k] = sos[k];
}

// One CTA (FIRLS_THREADS) per filter — scipy.signal.firls (_fir_filter_design.py:1130-1171) with
// desired = [0,0,1,1,0,0], weight = 1: q/b vectors, Q = toeplitz(q) + hankel(q) (n x n, n <= 64),
// Cholesky solve, symmetric tap assembly.
//
// The Cholesky is register tiled: the 64x64 matrix is cut into 16x16 blocks of 4x4 and the 136 lower-
// triangular blocks are dealt column-major to threads, each keeping its block in registers (block columns
// die left to right, so whole warps retire as the sweep advances).  Step k: the owners of column k publish the raw
// column to a double-buffered shared vector, the pivot's owner adds 1/sqrt(p) and 1/p, ONE barrier, then
// every live thread applies the rank-1 update to its registers.  No masking is needed: entries of finished
// rows/columns are dead, so updating them with stale values is harmless.  The right-hand side rides along
// as a 65th matrix row (16 extra threads), which folds the forward substitution into the same sweep; the
// scaled columns are kept (transposed) in shared memory for the backward substitution, done by warp 0 with
// shuffles.
constexpr int FIR_LDA = 65;
constexpr int FIRLS_THREADS = 160;   // 136 block owners + 16 rhs-row owners + 8 idle
__device__ void firls_design_block(double fs, int taps, double min_freq, double max_freq, double df,
                                   double* __restrict__ out, double* smem) {
  const int M = (taps - 1) / 2, n = M + 1;  // n unknowns (<= 64)
  double* q = smem;                 // [128]
  double* rhs = q + 128;            // [64]   b, later z, later the solution a
  double* colb = rhs + 64;          // [2][72] raw column k, [64] rhs-row entry, [65] 1/sqrt(p), [66] 1/p
  double* Lt = colb + 144;          // [64 * FIR_LDA]  Lt[k][i] = L[i][k]
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
  for (int i = t; i < 64; i += blockDim.x) rhs[i] = i < n ? fb[3] * np_sinc(fb[3] * i) - fb[2] * np_sinc(fb[2] * i) : 0.0;
  for (int i = t; i < 144; i += blockDim.x) colb[i] = 0.0;
  __syncthreads();
  // thread -> block: t < 136: lower-triangular block (bi, bj) in COLUMN-major order (block column bj is dead
  // once the sweep passes it, so the warps holding the early columns retire first); 136 <= t < 152: rhs row
  const bool owner = t < 136, rrow = t >= 136 && t < 152;
  int bi = 16, bj = rrow ? t - 136 : 0;
  if (owner) {
    int rem = t;
    bj = 0;
    while (rem >= 16 - bj) { rem -= 16 - bj; ++bj; }
    bi = bj + rem;
  }
  double a[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = 4 * bi + r, j = 4 * bj + c;
      if (owner) a[r][c] = (i < n && j < n) ? q[i > j ? i - j : j - i] + q[i + j] : (i == j ? 1.0 : 0.0);
      else if (rrow && r == 0) a[r][c] = rhs[j];
      else a[r][c] = 0.0;
    }
  const int nkb = (n + 3) >> 2;
  double* dinv = q;                 // q[] is dead once the blocks are loaded: reuse it for 1/L[k][k]
  __syncthreads();
  // publish column 0
  if (bj == 0) {
    double* col = colb;
    if (owner) {
      col[4 * bi + 0] = a[0][0]; col[4 * bi + 1] = a[1][0]; col[4 * bi + 2] = a[2][0]; col[4 * bi + 3] = a[3][0];
      if (bi == 0) { const double ri = rsqrt(a[0][0]); col[65] = ri; col[66] = ri * ri; }
    } else if (rrow) {
      col[64] = a[0][0];
    }
  }
  for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      const int k = 4 * kb + kc;
      const double* col = colb + (k & 1) * 72;
      double* ncol = colb + ((k + 1) & 1) * 72;
      __syncthreads();                       // column k (published during step k-1) is visible
      const double inv = col[65], invp = col[66];
      if (t < 64) Lt[k * FIR_LDA + t] = t >= k ? col[t] * inv : 0.0;     // L[t][k]
      if (t == 64) { rhs[k] = col[64] * inv; dinv[k] = inv; }             // z[k], 1/L[k][k]
      if (bj >= kb && (owner || rrow)) {       // finished block columns skip the update (whole warps retire)
        const double2 ja = *reinterpret_cast<const double2*>(col + 4 * bj), jb = *reinterpret_cast<const double2*>(col + 4 * bj + 2);
        const double lj[4] = {ja.x, ja.y, jb.x, jb.y};
        double li[4];
        if (owner) {
          const double2 ia = *reinterpret_cast<const double2*>(col + 4 * bi), ib = *reinterpret_cast<const double2*>(col + 4 * bi + 2);
          li[0] = -ia.x * invp; li[1] = -ia.y * invp; li[2] = -ib.x * invp; li[3] = -ib.y * invp;
        } else {
          li[0] = -col[64] * invp; li[1] = li[2] = li[3] = 0.0;
        }
        // look-ahead: bring column k+1 up to date first and publish it, so that the barrier of step k+1
        // overlaps the remaining 12 FMAs of step k
        const int nc = (kc + 1) & 3;           // column of k+1 inside its block (constant after unrolling)
        const int nkb2 = kc == 3 ? kb + 1 : kb;
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r][nc] = fma(li[r], lj[nc], a[r][nc]);
        if (kc != 3 && bj == nkb2 && k + 1 < 4 * nkb) {
          if (owner) {
            ncol[4 * bi + 0] = a[0][nc]; ncol[4 * bi + 1] = a[1][nc]; ncol[4 * bi + 2] = a[2][nc]; ncol[4 * bi + 3] = a[3][nc];
            if (bi == nkb2) { const double ri = rsqrt(a[nc][nc]); ncol[65] = ri; ncol[66] = ri * ri; }
          } else {
            ncol[64] = a[0][nc];
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c != nc) {
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r][c] = fma(li[r], lj[c], a[r][c]);
          }
      }
      // kc == 3: column k+1 lives in the NEXT block column, whose blocks are complete only after the full update
      if (kc == 3 && bj == kb + 1 && kb + 1 < nkb && (owner || rrow)) {
        if (owner) {
          ncol[4 * bi + 0] = a[0][0]; ncol[4 * bi + 1] = a[1][0]; ncol[4 * bi + 2] = a[2][0]; ncol[4 * bi + 3] = a[3][0];
          if (bi == kb + 1) { const double ri = rsqrt(a[0][0]); ncol[65] = ri; ncol[66] = ri * ri; }
        } else {
          ncol[64] = a[0][0];
        }
      }
    }
  }
  __syncthreads();
  // L^T x = z by warp 0: lane owns x[lane], x[lane + 32]
  if (t < 32) {
    double z0 = rhs[t], z1 = rhs[t + 32];
    for (int k = n - 1; k >= 0; --k) {
      const double zk = __shfl_sync(0xffffffffu, k < 32 ? z0 : z1, k & 31);
      const double xk = zk * dinv[k];
      if (t == (k & 31)) { if (k < 32) z0 = xk; else z1 = xk; }
      // z[i] -= L[k][i] * x[k] for i < k ;  L[k][i] = Lt[i][k]
      if (t < k) z0 = fma(-Lt[t * FIR_LDA + k], xk, z0);
      if (t + 32 < k) z1 = fma(-Lt[(t + 32) * FIR_LDA + k], xk, z1);
    }
    rhs[t] = z0; rhs[t + 32] = z1;
  }
  __syncthreads();
  // coeffs = [a[M..1], 2 a0, a[1..M]]
  for (int i = t; i < taps; i += blockDim.x) {
    const int d = i < M ? M - i : i - M;
    out[i] = d == 0 ? 2.0 * rhs[0] : rhs[d];
  }
}

__global__ void __launch_bounds__(FIRLS_THREADS) firls_from_fs_kernel(const double* __restrict__ fs, int taps, double min_freq,
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

__global__ void __launch_bounds__(FIRLS_THREADS) job_firls_kernel(const double* __restrict__ ring_t, bpv_window_params p,
                                                       double* __restrict__ out) {
  extern __shared__ double smem[];
  __shared__ int s_i[3];
  const int job = blockIdx.x;
  const int s = job / p.jobs_per_stream, j = job % p.jobs_per_stream;
  const double fs_s = job_fs_block(ring_t + (long long)s * p.cap, p.cap, p.window, p.head0 + (long long)j * p.head_step, s_i);
  firls_design_block(fs_s, p.fir_taps, p.min_freq, p.max_freq, p.fir_df, out + (long long)job * p.fir_taps, smem);
}

constexpr size_t FIRLS_SMEM = (128 + 64 + 144 + 64 * FIR_LDA) * sizeof(double);

int launch_job_butter(const double* ring_t, const bpv_window_params& p, double* sos_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_butter_kernel<<<(J + 63) / 64, 64, 0, st>>>(ring_t, p, sos_out);
  return check_launch("job_butter_kernel");
}

int launch_job_firls(const double* ring_t, const bpv_window_params& p, double* taps_out, cudaStream_t st) {
  const int J = p.S * p.jobs_per_stream;
  job_firls_kernel<<<J, FIRLS_THREADS, FIRLS_SMEM, st>>>(ring_t, p, taps_out);
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
  firls_from_fs_kernel<<<n, FIRLS_THREADS, FIRLS_SMEM, (cudaStream_t)stream>>>(fs, p->fir_taps, p->min_freq, p->max_freq, p->fir_df, taps_out);
  return check_launch("bpv_firls_design");
}
