// Register-tiled sliding dot product shared by the FIR filtfilt (F2) and the cross-correlation (F4):
//
//     acc[r] += sum_{k=0}^{K-1} c[k] * X[j0 + r - k],      r = 0 .. RT-1,   K % RT == 0
//
// A thread owns RT consecutive outputs, so consecutive taps reuse RT-1 of its RT operands: one new
// sample + one coefficient are loaded per RT FMAs (a plain "one output per thread" loop needs two
// shared-memory loads per FMA and is bound by shared-memory bandwidth, not by the FP64 pipe).
// X is stored "de-interleaved" — XT[(j % RT) * LD + j / RT] — so that the lanes of a warp, whose j0 are
// RT apart, read consecutive words (no bank conflicts).  j0 must be a multiple of RT and j0 - K >= 0.
// T = double (FIR filtfilt: every later bit-exact decision depends on it) or float (xcorr coarse pass).
#pragma once

namespace bpv {

template <int RT>
__device__ __forceinline__ int xt_index(int j, int LD) { return (j % RT) * LD + j / RT; }

// Optional [kb_begin, kb_end): only taps k in [kb_begin*RT, kb_end*RT) are applied (skips known-zero operands).
// LDC / KC != 0: leading dimension / padded tap count as compile-time constants.  With run-time values every tap block opens
// with a dependent chain (constant-bank load -> two IMADs -> LDS -> first FMA) and every operand load needs its own address
// computation: 29 integer instructions and ~17 % of the stall samples of the FIR loop (profiles/r2l_c2 source view).
// CVALID > 0 (run-time c_valid): c holds only c_valid coefficients (e.g. the taps as the design kernel left them in global
// memory) and the padding up to K is supplied as zeros here instead of by a zero-padded copy.
// C2: c is 16-byte aligned (T = double, RT even): the coefficients are fetched two at a time (LDS.128; the loads are
// warp-uniform, so this halves the coefficient wavefronts of a tap block).
template <int RT, typename T, int LDC = 0, int KC = 0, bool CGUARD = false, bool C2 = false>
__device__ __forceinline__ void corr_tile(T (&acc)[RT], const T* __restrict__ c, int K_rt,
                                          const T* __restrict__ XT, int LD_rt, int j0,
                                          int kb_begin = 0, int kb_end = 0x7fffffff, int c_valid = 0) {
  static_assert(!C2 || (sizeof(T) == 8 && RT % 2 == 0 && !CGUARD), "paired coefficient loads: double, even tile, no guard");
  const int LD = LDC ? LDC : LD_rt;
  const int K = KC ? KC : K_rt;
  const int col0 = j0 / RT;
  if (kb_end > K / RT) kb_end = K / RT;
  T w[RT];                            // w[s] = X[j] with j % RT == s, the RT samples under the current tap
#pragma unroll
  for (int s = 0; s < RT; ++s) w[s] = XT[s * LD + col0 - kb_begin];
  const T* cc = c + kb_begin * RT;
  const T* xn = XT + (col0 - kb_begin - 1);
#pragma unroll 1
  for (int kb = kb_begin; kb < kb_end; ++kb) {
#pragma unroll
    for (int kk = 0; kk < RT; ++kk) {
      T ck;
      if constexpr (C2) {
        const double2 cp = reinterpret_cast<const double2*>(cc)[kk >> 1];
        ck = (kk & 1) ? cp.y : cp.x;
      } else {
        ck = (!CGUARD || kb * RT + kk < c_valid) ? cc[kk] : (T)0;
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = fma(ck, w[(r - kk + RT) % RT], acc[r]);
      w[RT - 1 - kk] = xn[(RT - 1 - kk) * LD];      // X[j0 - k - 1] replaces X[j0 - k + RT - 1]
    }
    cc += RT;
    xn -= 1;
  }
}

// Circular variant for the full cross-correlation (F4): X holds the PERIODIC extension of the n-sample operand, so one
// sweep over the K taps visits, for output r of tile q, first the taps of lag li = RT*q - 1 + r (operand indices li - k >= 0)
// and — after the operand index wraps below zero, i.e. before tap k = RT*q + r — the taps of lag li + n.  The two lag
// ranges are complementary (li + 1 and n - 1 - li taps), so every lane does exactly K taps whatever its lags are: a
// linear sweep gives the centre lanes n taps and the outer lanes almost none, and the warp pays for the longest.
// At the wrap the accumulator of output r is moved to first[r] and cleared; on return acc[r] holds lag li + n.
// Same ascending tap order per lag as the linear form.  j0 % RT == 0, j0 - K >= 0, q < K / RT.
// LDC != 0: the leading dimension as a compile-time constant (the operand loads then use immediate offsets from one base
// register instead of an address computation per load).
template <int RT, int LDC, typename T>
__device__ __forceinline__ void corr_tile_wrap(T (&acc)[RT], T (&first)[RT], const T* __restrict__ c, int K,
                                               const T* __restrict__ XT, int LD_rt, int j0, int q) {
  const int LD = LDC ? LDC : LD_rt;
  const int col0 = j0 / RT;
  T w[RT];
#pragma unroll
  for (int s = 0; s < RT; ++s) w[s] = XT[s * LD + col0];
  const int nb = K / RT;
  for (int kb = 0; kb < nb; ++kb) {
    const T* cc = c + kb * RT;
    const T* xn = XT + (col0 - kb - 1);
    const bool wrap = kb == q;
#pragma unroll
    for (int kk = 0; kk < RT; ++kk) {
      if (wrap) { first[kk] = acc[kk]; acc[kk] = (T)0; }      // output kk wraps before tap RT*q + kk
      const T ck = cc[kk];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = fma(ck, w[(r - kk + RT) % RT], acc[r]);
      w[RT - 1 - kk] = xn[(RT - 1 - kk) * LD];
    }
  }
}

// Packed-FMA form of corr_tile_wrap for float operands (sm_100a `fma.rn.f32x2`, SASS FFMA2: two float FMAs per issued
// instruction with a scalar-broadcast first operand and a per-operand half swap).  The RT outputs of a lane are held as
// RT/2 register pairs (acc[i], acc[i + RT/2]) and its RT operand samples as pairs (w[s], w[s + RT/2]): under tap kk
// output i needs w[(i - kk) mod RT] and output i + RT/2 needs the sample RT/2 further on — the two halves of ONE pair,
// swapped when (i - kk) mod RT >= RT/2.  Every lane-level FMA is the one corr_tile_wrap issues, in the same order, so the
// results are identical bit for bit; the loop body is RT/2 FFMA2 per tap instead of RT FFMA (the kernel is issue bound).
// c must be 8-byte aligned; RT even.
template <int RT, int LDC>
__device__ __forceinline__ void corr_tile_wrap_f32x2(float (&acc)[RT], float (&first)[RT], const float* __restrict__ c, int K,
                                                     const float* __restrict__ XT, int LD_rt, int j0, int q) {
  static_assert(RT % 2 == 0, "pairs of outputs");
  constexpr int H = RT / 2;
  const int LD = LDC ? LDC : LD_rt;
  const int col0 = j0 / RT;
  float2 A[H], P[H];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    A[i] = make_float2(acc[i], acc[i + H]);
    P[i] = make_float2(XT[i * LD + col0], XT[(i + H) * LD + col0]);
  }
  const int nb = K / RT;
  for (int kb = 0; kb < nb; ++kb) {
    const float2* cc = reinterpret_cast<const float2*>(c + kb * RT);
    const float* xn = XT + (col0 - kb - 1);
    const bool wrap = kb == q;
#pragma unroll
    for (int kk = 0; kk < RT; ++kk) {
      if (wrap) {                                             // output kk wraps before tap RT*q + kk
        if (kk < H) { first[kk] = A[kk].x; A[kk].x = 0.f; }
        else { first[kk] = A[kk - H].y; A[kk - H].y = 0.f; }
      }
      const float2 cp = cc[kk / 2];
      const float ck = (kk & 1) ? cp.y : cp.x;
      const float2 cb = make_float2(ck, ck);
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const int lo = (i - kk + RT) % RT;                    // sample under output i; output i + H sees (lo + H) % RT
        const float2 op = lo < H ? P[lo] : make_float2(P[lo - H].y, P[lo - H].x);
        A[i] = __ffma2_rn(cb, op, A[i]);
      }
      const int t = RT - 1 - kk;                              // X[j0 - k - 1] replaces X[j0 - k + RT - 1]
      const float xv = xn[t * LD];
      if (t < H) P[t].x = xv; else P[t - H].y = xv;
    }
  }
#pragma unroll
  for (int i = 0; i < H; ++i) { acc[i] = A[i].x; acc[i + H] = A[i].y; }
}

}  // namespace bpv
