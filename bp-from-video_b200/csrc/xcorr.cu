// F4(b) — pairwise-ROI cross-correlation and PTT lag search (SignalProcessor.correlate_signal_pair /
// correlate_signals and the get_peaks() on sg_corr: signal_processor.py:280-299, 312).
// One CTA per (window job, ROI pair); the jointly valid samples of both signals sit in shared memory;
// each thread owns lags k, k+blockDim, ... of the 2n-1 full-mode lags (float64 dot products, the
// "direct" method scipy.signal.correlate picks for these sizes); normalisation by max(a.a, b.b, a.b);
// first-max argmax over every finite lag fused in.
#include "filters.cuh"
#include "corr_tile.cuh"

namespace bpv {

constexpr float XC_DELTA = 2.0e-4f;   // candidate band below the fp32 maximum (normalised correlation, |c| <~ 1)

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
  // shared memory: doubles xa[W] | a64[W] | b64[W] ; floats cv[2W] | c[Kw] reversed b, zero padded | XT[8*LD] a, de-interleaved
  constexpr int RT = 8;
  const int Kw = (W + RT - 1) / RT * RT;
  const int LDw = (Kw + 3 * W + 16) / RT + 1;
  double* xa = sm;             // [W] timestamps of signal a
  double* a64 = xa + W;        // [W] jointly valid samples, float64 (normalisation + refinement)
  double* b64 = a64 + W;       // [W]
  float* cv = reinterpret_cast<float*>(b64 + W);   // [2W] coarse correlation values
  float* c = cv + 2 * W;       // [Kw]
  float* XT = c + Kw;          // [RT * LDw]
  for (int i = tid; i < Kw; i += blockDim.x) c[i] = 0.f;
  for (int i = tid; i < RT * LDw; i += blockDim.x) XT[i] = 0.f;
  __syncthreads();
  // stage the three windows (coalesced), then warp 0 compacts the jointly valid samples in place:
  // valid = a.w & b.w (finite in both)
  for (int k = tid; k < W; k += blockDim.x) { xa[k] = xa_g[k]; a64[k] = ya_g[k]; b64[k] = yb_g[k]; }
  __syncthreads();
  if (tid < 32) {
    int cnt = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int k0 = 0; k0 < W; k0 += 32) {
      const int k = k0 + lane;
      double va = nan_f64(), vb = nan_f64(), vx = nan_f64();
      if (k < W) { va = a64[k]; vb = b64[k]; vx = xa[k]; }
      const bool ok = isfinite(va) && isfinite(vb);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      __syncwarp();
      if (ok) {
        const int i = cnt + __popc(bal & lt);
        a64[i] = va; b64[i] = vb; xa[i] = vx;
      }
      __syncwarp();
      cnt += __popc(bal);
    }
    if (lane == 0) s_n = cnt;
  }
  __syncthreads();
  const int n = s_n;
  if (n < 2) {                                  // guard signal_processor.py:284 -> empty
    if (tid == 0) { num_lags[jp] = 0; lag_idx[jp] = -1; lag_sec[jp] = nan_f64(); lag_corr[jp] = nan_f64(); }
    return;
  }
  const int K = (n + RT - 1) / RT * RT;
  double daa = 0, dbb = 0, dab = 0, amax = 0, bmax = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double va = a64[i], vb = b64[i];
    daa = fma(va, va, daa); dbb = fma(vb, vb, dbb); dab = fma(va, vb, dab);
    amax = fmax(amax, fabs(va)); bmax = fmax(bmax, fabs(vb));
  }
  daa = block_sum(daa, s_val); dbb = block_sum(dbb, s_val); dab = block_sum(dab, s_val);
  const double den = fmax(fmax(daa, dbb), dab);
  // fp32 operands scaled to O(1) so tiny band-passed signals neither underflow nor lose bits
  for (int o = 16; o > 0; o >>= 1) { amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o)); bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o)); }
  __syncthreads();
  if ((tid & 31) == 0) { s_val[tid >> 5] = amax; s_val[8 + (tid >> 5)] = bmax; }
  __syncthreads();
  for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) { amax = fmax(amax, s_val[w2]); bmax = fmax(bmax, s_val[8 + w2]); }
  __syncthreads();
  const double sa = amax > 0 ? 1.0 / amax : 1.0, sb = bmax > 0 ? 1.0 / bmax : 1.0;
  for (int i = tid; i < n; i += blockDim.x) {
    XT[xt_index<RT>(i + n - 1 + K, LDw)] = (float)(a64[i] * sa);     // storage index = u + K
    c[n - 1 - i] = (float)(b64[i] * sb);
  }
  __syncthreads();
  const float unscale = (float)(1.0 / (sa * sb * den));
  const int L = 2 * n - 1;
  const long long ob = jp * (2LL * W - 1);
  // COARSE pass in fp32:  corr[li] = sum_l a[l + k] b[l], k = li - (n-1)  ==  sum_m c[m] X[j - m] with j = li + n - 1
  // (corr_tile.cuh); X is non-zero only on u in [n-1, 2n-2], so each tile applies just the taps that can touch it.
  const int jbase = (n - 1 + K) / RT * RT;      // storage index of the tile that holds li = 0
  for (int J0 = jbase + RT * tid; J0 <= (L - 1) + (n - 1) + K; J0 += RT * blockDim.x) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.f;
    const int u0 = J0 - K;                       // logical index of the tile's first output operand
    int kb_lo = (u0 - (2 * n - 2)) / RT; if (kb_lo < 0) kb_lo = 0;
    int kb_hi = (u0 + RT - 1 - (n - 1)) / RT + 1; if (kb_hi < 0) kb_hi = 0;
    corr_tile<RT, float>(acc, c, K, XT, LDw, J0, kb_lo, kb_hi);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int li = u0 + r - (n - 1);
      if (li >= 0 && li < L) {
        const float cc = acc[r] * unscale;
        cv[li] = cc;
        if (corr_val) {
          const int k = li - (n - 1), ak = k < 0 ? -k : k;
          const double lag = (xa[n - 1] - xa[n - 1 - ak]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
          corr_lag[ob + li] = (float)lag;
          corr_val[ob + li] = cc;
        }
      }
    }
  }
  __syncthreads();
  // PEAK in float64: every lag whose coarse value is within XC_DELTA of the coarse maximum is re-evaluated as a
  // float64 dot product; the first maximum among those decides (Signal.get_peak after the range reset).
  float cmax = -INFINITY; int cnt = 0;
  for (int li = tid; li < L; li += blockDim.x) { const float v = cv[li]; if (isfinite(v)) { ++cnt; cmax = fmaxf(cmax, v); } }
  for (int o = 16; o > 0; o >>= 1) { cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  {
    const int wid = tid >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { s_val[wid] = cmax; s_idx[wid] = cnt; }
    __syncthreads();
    cmax = s_val[0]; cnt = s_idx[0];
    for (int w2 = 1; w2 < nw; ++w2) { cmax = fmaxf(cmax, (float)s_val[w2]); cnt += s_idx[w2]; }
    __syncthreads();
  }
  double bv = -INFINITY; int bi = 0x7fffffff;
  for (int li = tid; li < L; li += blockDim.x) {
    const float v = cv[li];
    if (cnt >= 2 ? (isfinite(v) && v >= cmax - XC_DELTA) : true) {
      const int k = li - (n - 1);
      const int l0 = k < 0 ? -k : 0, l1 = k > 0 ? n - k : n;
      double acc = 0.0;
      for (int l = l0; l < l1; ++l) acc = fma(a64[l + k], b64[l], acc);
      const double cc = acc / den;
      if (isfinite(cc) && (cc > bv || (cc == bv && li < bi))) { bv = cc; bi = li; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  {
    const int wid = tid >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w2 = 1; w2 < nw; ++w2)
        if (s_val[w2] > bv || (s_val[w2] == bv && s_idx[w2] < bi)) { bv = s_val[w2]; bi = s_idx[w2]; }
      // finite-lag count in float64 terms: den == 0 or non-finite makes every lag non-finite (NaN / inf)
      const bool any = isfinite(den) && den != 0.0 && bi != 0x7fffffff;
      num_lags[jp] = L;
      if (any && L >= 2) {
        const int k = bi - (n - 1), ak = k < 0 ? -k : k;
        lag_idx[jp] = bi;
        lag_sec[jp] = (xa[n - 1] - xa[n - 1 - ak]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
        lag_corr[jp] = bv;
      } else { lag_idx[jp] = -1; lag_sec[jp] = nan_f64(); lag_corr[jp] = nan_f64(); }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Warp-per-pair version (default): the CTA-per-pair kernel above spends most of its time in __syncthreads between short
// phases (staging, single-warp compaction, five block reductions, peak search: the tile phase is 18 % of its samples).
// Here one warp owns a (job, pair): every phase is warp-local (ballot compaction, shuffle reductions), each lane owns the
// XW_RT = 10 lag pairs (li, li + n) of one tile, so a 300-sample window (599 lags) is a single round of the register-tiled
// circular sliding dot product (2 LDS + 2 predicated moves per 10 FFMA, n taps per lane), and the float64 re-evaluation of
// the candidate lags is a warp-cooperative dot product.
// smem per warp: doubles a64[W] | b64[W]; u16 pos[W]; floats cv[2W] (only when a lane owns more than one tile, W > 320) | c[K] | XT[RT * LD]
// ---------------------------------------------------------------------------------------------
constexpr int XW_RT = 10;
struct XwLayout { int b64, pos, cv, c, xt, total, K, LD, one_round; };
__host__ __device__ inline XwLayout xw_layout(int W, bool force_cv = false) {
  XwLayout L;
  L.K = (W + XW_RT - 1) / XW_RT * XW_RT;                 // taps, padded to the tile
  L.LD = (2 * (L.K / XW_RT) + 2) | 1;                    // columns of the de-interleaved periodic operand (odd: no bank conflicts)
  int o = W * 8;
  L.b64 = o; o += W * 8;
  L.pos = o; o += (W * 2 + 15) / 16 * 16;
  L.one_round = !force_cv && L.K / XW_RT <= 32;                       // every lane owns at most one tile: coarse values stay in registers
  L.cv = o; o += L.one_round ? 0 : 2 * W * 4;
  L.c = o; o += L.K * 4;
  L.xt = o; o += XW_RT * L.LD * 4;
  L.total = (o + 15) / 16 * 16;
  return L;
}

// COARSE pass layout (corr_tile.cuh, corr_tile_wrap):  corr[li] = sum_m c[m] a[li - m]  (c[m] = b[n-1-m], 0 <= li - m < n).
// Lags li and li + n use complementary tap ranges, so a lane owns the lag PAIRS li = RT*tile - 1 + r, r = 0..RT-1 and sweeps the
// K taps once over the periodic extension AA[t] = a[t mod n], t in [-n, n), stored at X[t + K + RT + 1]: n multiply-adds per
// output pair for every lane (a 300-sample window: 30 lanes x 10 pairs, one round), where one tile of consecutive lags per
// lane made the warp wait for the centre lanes' n taps per lag.
// LDC = the operand buffer's leading dimension when it is one of the specialised window sizes (0 = run-time value)
// ONE = every lane owns at most one tile (windows up to 320 samples): the 2 x RT coarse values of a lane never leave its
// registers — the candidate lags are flagged from them as a 20-bit mask per lane — so the plan has no cv[] array (9.1 instead of
// 11.5 KB per warp at W = 300) and the peak search does not re-read 2n - 1 values from shared memory.  Launch bound 5 CTAs per SM:
// 88 registers keep the packed tile loop free of spills (ptxas settles on 72 without it and pays 60 register moves per 10 taps).
template <int LDC, bool ONE>
__global__ void __launch_bounds__(128, ONE ? 5 : 1) xcorr_warp_kernel(const double* __restrict__ proc_x, const double* __restrict__ proc_y,
                                                         const bpv_window_params p, const XwLayout Lw, long long npairs,
                                                         float* __restrict__ corr_lag, float* __restrict__ corr_val,
                                                         int32_t* __restrict__ num_lags, int32_t* __restrict__ lag_idx,
                                                         double* __restrict__ lag_sec, double* __restrict__ lag_corr) {
  extern __shared__ __align__(16) unsigned char xsm[];
  constexpr int RT = XW_RT;
  const int W = p.window, R = p.R, P = R * (R - 1) / 2, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long jp = (long long)blockIdx.x * (blockDim.x >> 5) + wid;   // job * P + pair
  if (jp >= npairs) return;
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
  unsigned char* sm = xsm + (size_t)wid * Lw.total;
  double* a64 = reinterpret_cast<double*>(sm);
  double* b64 = reinterpret_cast<double*>(sm + Lw.b64);
  unsigned short* pos = reinterpret_cast<unsigned short*>(sm + Lw.pos);
  float* cv = reinterpret_cast<float*>(sm + Lw.cv);
  float* c = reinterpret_cast<float*>(sm + Lw.c);
  float* XT = reinterpret_cast<float*>(sm + Lw.xt);
  const int LD = LDC ? LDC : Lw.LD;
  // stage both windows (independent coalesced loads)
  bool allf = true;
  for (int k = lane; k < W; k += 32) {
    const double va = ya_g[k], vb = yb_g[k];
    a64[k] = va; b64[k] = vb; pos[k] = (unsigned short)k;
    allf &= isfinite(va) && isfinite(vb);
  }
  __syncwarp();
  // jointly valid samples (valid = a.w & b.w), compacted in place; a pair of windows without holes already is (one vote
  // instead of W / 32 ballot rounds, positions = identity)
  int n = 0;
  if (__all_sync(0xffffffffu, allf)) {
    n = W;
  } else {
    const unsigned lt = (1u << lane) - 1u;
    for (int k0 = 0; k0 < W; k0 += 32) {
      const int k = k0 + lane;
      double va = nan_f64(), vb = nan_f64();
      if (k < W) { va = a64[k]; vb = b64[k]; }
      const bool ok = isfinite(va) && isfinite(vb);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      __syncwarp();
      if (ok) { const int i = n + __popc(bal & lt); a64[i] = va; b64[i] = vb; pos[i] = (unsigned short)k; }
      __syncwarp();
      n += __popc(bal);
    }
  }
  if (n < 2) {                                  // guard signal_processor.py:284 -> empty
    if (lane == 0) { num_lags[jp] = 0; lag_idx[jp] = -1; lag_sec[jp] = nan_f64(); lag_corr[jp] = nan_f64(); }
    return;
  }
  const int K = (n + RT - 1) / RT * RT;         // <= Lw.K
  const int OFF = K + RT + 1;                   // storage index of AA[0]
  double daa = 0, dbb = 0, dab = 0, amax = 0, bmax = 0;
  for (int i = lane; i < n; i += 32) {
    const double va = a64[i], vb = b64[i];
    daa = fma(va, va, daa); dbb = fma(vb, vb, dbb); dab = fma(va, vb, dab);
    amax = fmax(amax, fabs(va)); bmax = fmax(bmax, fabs(vb));
  }
  daa = warp_sum(daa); dbb = warp_sum(dbb); dab = warp_sum(dab);
  for (int o = 16; o > 0; o >>= 1) { amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o)); bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o)); }
  const double den = fmax(fmax(daa, dbb), dab);
  // fp32 operands scaled to O(1) so tiny band-passed signals neither underflow nor lose bits.  The operand buffer is
  // written completely (every storage index a tile can touch): periodic extension inside [-n, n), zeros outside.
  const double sa = amax > 0 ? 1.0 / amax : 1.0, sb = bmax > 0 ? 1.0 / bmax : 1.0;
  const int jtot = 2 * K + 2 * RT;              // storage indices [0, jtot): t = j - OFF in [-K - RT - 1, K + RT - 2]
  // each sample is converted once and stored at its two periods (t = i and t = i - n); the few slots outside [-n, n) are zeroed
  for (int i = lane; i < n; i += 32) {
    const float v = (float)(a64[i] * sa);
    XT[xt_index<RT>(i + OFF, LD)] = v;
    XT[xt_index<RT>(i + OFF - n, LD)] = v;
  }
  for (int j = lane; j < OFF - n; j += 32) XT[xt_index<RT>(j, LD)] = 0.f;
  for (int j = OFF + n + lane; j < jtot; j += 32) XT[xt_index<RT>(j, LD)] = 0.f;
  for (int m = lane; m < K; m += 32) c[m] = m < n ? (float)(b64[n - 1 - m] * sb) : 0.f;
  __syncwarp();
  const float unscale = (float)(1.0 / (sa * sb * den));
  const int L = 2 * n - 1;
  const long long ob = jp * (2LL * W - 1);
  const double x_last = xa_g[pos[n - 1]];
  float cmax = -INFINITY; int cnt = 0;          // coarse maximum / finite count, gathered while the tiles are written
  const int tiles = K / RT;
  unsigned okm = 0;                             // ONE: bit r = first lag of pair r is a lag of this window, bit RT + r = second lag
  float acc[RT], first[RT];                     // ONE: after the tile, the lane's scaled coarse values (second | first lags)
  for (int tile = lane; tile < tiles; tile += 32) {
#pragma unroll
    for (int r = 0; r < RT; ++r) { acc[r] = 0.f; first[r] = 0.f; }
    corr_tile_wrap_f32x2<RT, LDC>(acc, first, c, K, XT, LD, RT * tile + K + RT, tile);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int la = RT * tile - 1 + r;          // first lag of the pair; the second is la + n
      first[r] *= unscale; acc[r] *= unscale;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int li = h ? la + n : la;
        const bool ok = h ? (la <= n - 2) : (la >= 0 && la <= n - 2);
        if (ok) {
          const float cc = h ? acc[r] : first[r];
          if (ONE) okm |= 1u << (h * RT + r); else cv[li] = cc;
          if (isfinite(cc)) { ++cnt; cmax = fmaxf(cmax, cc); }
          if (corr_val) {
            const int k = li - (n - 1), ak = k < 0 ? -k : k;
            const double lag = (x_last - xa_g[pos[n - 1 - ak]]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
            corr_lag[ob + li] = (float)lag;
            corr_val[ob + li] = cc;
          }
        }
      }
    }
    if (ONE) break;                             // tiles <= 32: one tile per lane
  }
  __syncwarp();
  // PEAK in float64: every lag whose coarse value is within XC_DELTA of the coarse maximum is re-evaluated as a
  // float64 dot product (warp cooperative); the first maximum among those decides (Signal.get_peak after the range reset).
  for (int o = 16; o > 0; o >>= 1) { cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  double bv = -INFINITY; int bi = 0x7fffffff;
  const float thr = cmax - XC_DELTA;
  auto refine = [&](int lc) {                    // whole warp: lag index lc -> float64 correlation, first-max update
    const int k = lc - (n - 1);
    const int l0 = k < 0 ? -k : 0, l1 = k > 0 ? n - k : n;
    double acc = 0.0;
    for (int l = l0 + lane; l < l1; l += 32) acc = fma(a64[l + k], b64[l], acc);
    acc = warp_sum(acc);
    const double cc = acc / den;
    if (isfinite(cc) && (cc > bv || (cc == bv && lc < bi))) { bv = cc; bi = lc; }
  };
  if (ONE) {
    unsigned cand = 0;
#pragma unroll
    for (int b = 0; b < 2 * RT; ++b) {
      const float v = b < RT ? first[b] : acc[b - RT];
      if (((okm >> b) & 1u) && (cnt >= 2 ? (isfinite(v) && v >= thr) : true)) cand |= 1u << b;
    }
    unsigned m = __ballot_sync(0xffffffffu, cand != 0);
    while (m) {                                  // any order: the update keeps the largest value, ties to the smallest lag
      const int src = __ffs(m) - 1;
      m &= m - 1;
      unsigned f = __shfl_sync(0xffffffffu, cand, src);
      while (f) {
        const int b = __ffs(f) - 1;
        f &= f - 1;
        const int la = RT * src - 1 + (b < RT ? b : b - RT);
        refine(b < RT ? la : la + n);
      }
    }
  } else {
    for (int li0 = 0; li0 < L; li0 += 128) {       // 4 consecutive lags per lane per step
      const int lb = li0 + 4 * lane;
      unsigned flags = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int li = lb + e;
        if (li < L) { const float v = cv[li]; if (cnt >= 2 ? (isfinite(v) && v >= thr) : true) flags |= 1u << e; }
      }
      unsigned m = __ballot_sync(0xffffffffu, flags != 0);
      while (m) {                                  // lanes in increasing lag order, lags of a lane in increasing order
        const int src = __ffs(m) - 1;
        m &= m - 1;
        unsigned f = __shfl_sync(0xffffffffu, flags, src);
        while (f) {
          const int lc = li0 + 4 * src + __ffs(f) - 1;
          f &= f - 1;
          refine(lc);
        }
      }
    }
  }
  if (lane == 0) {
    // finite-lag count in float64 terms: den == 0 or non-finite makes every lag non-finite (NaN / inf)
    const bool any = isfinite(den) && den != 0.0 && bi != 0x7fffffff;
    num_lags[jp] = L;
    if (any && L >= 2) {
      const int k = bi - (n - 1), ak = k < 0 ? -k : k;
      lag_idx[jp] = bi;
      lag_sec[jp] = (x_last - xa_g[pos[n - 1 - ak]]) * (k > 0 ? 1.0 : (k < 0 ? -1.0 : 0.0));
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
  {  // warp-per-pair kernel whenever one pair fits the shared memory of a CTA
    // BPV_XC_ONE=0 (measurement switch): coarse values through the cv[] array in shared memory, as before round 2's last change
    static const bool force_cv = [] { const char* e = getenv("BPV_XC_ONE"); return e && e[0] == '0'; }();
    const XwLayout Lw = xw_layout(W, force_cv);
    int wpb = (200 * 1024) / Lw.total;
    if (wpb > 4) wpb = 4;
    if (wpb >= 1) {
      const size_t smw = (size_t)wpb * Lw.total;
      const unsigned grid = (unsigned)((n + wpb - 1) / wpb);
#define BPV_XW(LDC, ONE)                                                                                              \
  do {                                                                                                                \
    if (int rc = ensure_dyn_smem((const void*)xcorr_warp_kernel<LDC, ONE>, smw)) return rc;                           \
    xcorr_warp_kernel<LDC, ONE><<<grid, wpb * 32, smw, (cudaStream_t)stream>>>(proc_x, proc_y, *p, Lw, n, corr_lag,  \
                                                                              corr_val, num_lags, lag_idx, lag_sec, lag_corr); \
  } while (0)
      if (Lw.LD == 63 && Lw.one_round) BPV_XW(63, true);      // windows of 291..300 samples (the 10 s window at 30 fps)
      else if (Lw.LD == 53 && Lw.one_round) BPV_XW(53, true); // 241..250 (the reference's default signal_max_samples)
      else if (Lw.LD == 63) BPV_XW(63, false);
      else if (Lw.one_round) BPV_XW(0, true);
      else BPV_XW(0, false);
#undef BPV_XW
      return check_launch("bpv_window_xcorr");
    }
  }
  const int Kw = (W + 7) / 8 * 8;
  const size_t smem = (size_t)(3 * W) * sizeof(double) + (size_t)(2 * W + Kw + 8 * ((Kw + 3 * W + 16) / 8 + 1)) * sizeof(float);
  BPV_REQUIRE(smem <= 200 * 1024, BPV_E_TOO_LARGE, "bpv_window_xcorr: window %d too large for shared memory", W);
  if (int rc = ensure_dyn_smem((const void*)xcorr_kernel, smem)) return rc;
  xcorr_kernel<<<(unsigned)n, 128, smem, (cudaStream_t)stream>>>(proc_x, proc_y, *p, corr_lag, corr_val, num_lags, lag_idx, lag_sec, lag_corr);
  return check_launch("bpv_window_xcorr");
}
