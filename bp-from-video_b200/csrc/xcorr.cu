// F4(b) — pairwise-ROI cross-correlation and PTT lag search (SignalProcessor.correlate_signal_pair /
// correlate_signals and the get_peaks() on sg_corr: signal_processor.py:280-299, 312).
// One CTA per (window job, ROI pair); the jointly valid samples of both signals sit in shared memory;
// each thread owns lags k, k+blockDim, ... of the 2n-1 full-mode lags (float64 dot products, the
// "direct" method scipy.signal.correlate picks for these sizes); normalisation by max(a.a, b.b, a.b);
// first-max argmax over every finite lag fused in.
#include "filters.cuh"

namespace bpv {

__global__ void __launch_bounds__(128) xcorr_kernel(const double* __restrict__ proc_x, const double* __restrict__ proc_y,
                                                    const bpv_window_params p, float* __restrict__ corr_lag,
                                                    float* __restrict__ corr_val, int32_t* __restrict__ num_lags,
                                                    int32_t* __restrict__ lag_idx, double* __restrict__ lag_sec,
                                                    double* __restrict__ lag_corr) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_val[33];
  __shared__ int s_idx[64];
  __shared__ int s_n;
  const int W = p.window, R = p.R, P = R * (R - 1) / 2, tid = threadIdx.x, lane = tid & 31;
  const long long jp = blockIdx.x;             // job * P + pair
  const long long job = jp / P;
  int pr = (int)(jp % P), ra = 0, rb = 1;
  for (ra = 0; ra < R - 1; ++ra) {             // itertools.combinations order
    const int cnt = R - 1 - ra;
    if (pr < cnt) { rb = ra + 1 + pr; break; }
    pr -= cnt;
  }
  const double* xa_g = proc_x + (job * R + ra) * W;
  const double* ya_g = proc_y + (job * R + ra) * W;
  const double* yb_g = proc_y + (job * R + rb) * W;
  double* a = sm;            // [W]
  double* b = a + W;         // [W]
  double* xa = b + W;        // [W]
  double* cv = xa + W;       // [2W-1] correlation values (for the argmax)
  // valid = a.w & b.w  (finite in both)
  if (tid < 32) {
    int n = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int k0 = 0; k0 < W; k0 += 32) {
      const int k = k0 + lane;
      double va = nan_f64(), vb = nan_f64(), vx = nan_f64();
      if (k < W) { va = ya_g[k]; vb = yb_g[k]; vx = xa_g[k]; }
      const bool ok = isfinite(va) && isfinite(vb);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) { const int i = n + __popc(bal & lt); a[i] = va; b[i] = vb; xa[i] = vx; }
      n += __popc(bal);
    }
    if (lane == 0) s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  if (n < 2) {                                  // guard signal_processor.py:284 -> empty
    if (tid == 0) { num_lags[jp] = 0; lag_idx[jp] = -1; lag_sec[jp] = nan_f64(); lag_corr[jp] = nan_f64(); }
    return;
  }
  double daa = 0, dbb = 0, dab = 0;
  for (int i = tid; i < n; i += blockDim.x) { daa = fma(a[i], a[i], daa); dbb = fma(b[i], b[i], dbb); dab = fma(a[i], b[i], dab); }
  daa = block_sum(daa, s_val); dbb = block_sum(dbb, s_val); dab = block_sum(dab, s_val);
  const double den = fmax(fmax(daa, dbb), dab);
  const int L = 2 * n - 1;
  const long long ob = jp * (2LL * W - 1);
  for (int li = tid; li < L; li += blockDim.x) {
    const int k = li - (n - 1);                // corr[k + n - 1] = sum_l a[l + k] * b[l]
    const int l0 = k < 0 ? -k : 0, l1 = k > 0 ? n - k : n;
    double acc = 0.0;
    for (int l = l0; l < l1; ++l) acc = fma(a[l + k], b[l], acc);
    const double c = acc / den;
    cv[li] = c;
    if (corr_val) {
      const int ak = k < 0 ? -k : k;
      const double lag = (xa[n - 1] - xa[n - 1 - ak]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
      corr_lag[ob + li] = (float)lag;
      corr_val[ob + li] = (float)c;
    }
  }
  __syncthreads();
  // first-max over finite correlation values (Signal.get_peak after the range reset)
  double bv = -INFINITY; int bi = 0x7fffffff, cnt = 0;
  for (int li = tid; li < L; li += blockDim.x) {
    const double v = cv[li];
    if (isfinite(v)) { ++cnt; if (v > bv || (v == bv && li < bi)) { bv = v; bi = li; } }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int wid = tid >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; s_idx[32 + wid] = cnt; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nw; ++w) {
      if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
      cnt += s_idx[32 + w];
    }
    num_lags[jp] = L;
    if (cnt >= 2) {
      const int k = bi - (n - 1), ak = k < 0 ? -k : k;
      lag_idx[jp] = bi;
      lag_sec[jp] = (xa[n - 1] - xa[n - 1 - ak]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
      lag_corr[jp] = bv;
    } else { lag_idx[jp] = -1; lag_sec[jp] = nan_f64(); lag_corr[jp] = nan_f64(); }
  }
}

}  // namespace bpv

extern "C" int bpv_window_xcorr(const double* proc_x, const double* proc_y, const bpv_window_params* p,
                                float* corr_lag, float* corr_val, int32_t* num_lags,
                                int32_t* lag_idx, double* lag_sec, double* lag_corr, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(p && proc_x && proc_y && num_lags && lag_idx && lag_sec && lag_corr, BPV_E_INVALID, "bpv_window_xcorr: NULL pointer");
  BPV_REQUIRE((corr_lag == nullptr) == (corr_val == nullptr), BPV_E_INVALID, "bpv_window_xcorr: corr_lag/corr_val must both be set or both NULL");
  const int W = p->window, P = p->R * (p->R - 1) / 2;
  if (P == 0) return 0;
  const long long n = (long long)p->S * p->jobs_per_stream * P;
  BPV_REQUIRE(W > 0 && n > 0, BPV_E_INVALID, "bpv_window_xcorr: bad sizes");
  const size_t smem = (size_t)(5 * W) * sizeof(double);
  BPV_REQUIRE(smem <= 200 * 1024, BPV_E_TOO_LARGE, "bpv_window_xcorr: window %d too large for shared memory", W);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(xcorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  xcorr_kernel<<<(unsigned)n, 128, smem, (cudaStream_t)stream>>>(proc_x, proc_y, *p, corr_lag, corr_val, num_lags, lag_idx, lag_sec, lag_corr);
  return check_launch("bpv_window_xcorr");
}
