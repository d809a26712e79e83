// F3 + F4 in ONE grid — transform_signals + get_peaks (signal_processor.py:248-277, 310) with PGRAM_WELCH and
// correlate_signals + get_peaks (:280-299, 312) of the same window jobs.
//
// Both kernels only read the processed windows, and they are bound by different things: the Welch kernel by shared-memory
// wavefronts of its float64 FFT (issue slots 44 % used), the cross-correlation by FP32 issue (shared memory 37 %).  Launched
// on two streams they barely mix — the block scheduler places the first kernel's CTAs until none are left, so the
// second runs in its tail (profiles/r4c: Welch 110 us as if alone, xcorr's remaining 118 of its 130 us behind it).  Here CTAs
// of the two roles are INTERLEAVED in one grid in the ratio of their counts (R = 2: two Welch CTAs of four signals per xcorr
// CTA of four pairs), so every SM holds a mix of both for the whole launch.  The role of a CTA is a function of its index;
// each role runs the unmodified body of its stand-alone kernel (welch.cuh, xcorr_warp.cuh): results are identical bit for bit.
//
// MEASURED (profiles/r4g_bench_c2_fused.json against r4g_bench_c2.json, 16 384 window jobs): 272.9 us for the fused grid against
// 227.5 us for the two kernels on two streams (240 one after the other) — the mix is SLOWER.  What it costs: both roles at the
// larger role's footprint (96 registers, 36 KB), 41 instead of 26 register moves per 10 taps of the packed xcorr tile under
// the joint register allocation, and two large unrolled bodies (6.6 k instructions) resident on every SM at once.  Not used by
// default (engine overlap bit 4); kept as a tested experiment.
#include "welch.cuh"
#include "xcorr_warp.cuh"

namespace bpv {

struct WelchArgs {
  const double* proc_x; const double* proc_y; int max_bins; long long nsig;
  float* spec_f; float* spec_mag; int32_t* num_bins; int32_t* peak_idx; double* peak_freq; double* peak_mag;
};
struct XcorrArgs {
  const double* proc_x; const double* proc_y; XwLayout Lw; long long npairs;
  float* corr_lag; float* corr_val; int32_t* num_lags; int32_t* lag_idx; double* lag_sec; double* lag_corr;
};

template <int LDC>
__global__ void __launch_bounds__(128, 5) welch_xcorr_kernel(const bpv_window_params p, const WelchArgs wa, const XcorrArgs xa,
                                                             unsigned n_welch, unsigned n_xcorr) {
  extern __shared__ __align__(16) unsigned char fsm[];
  // CTA b is the x-th xcorr CTA when floor((b + 1) n_xcorr / total) > floor(b n_xcorr / total) = x, else Welch CTA b - x
  const unsigned b = blockIdx.x;
  const unsigned long long tot = (unsigned long long)n_welch + n_xcorr;
  const unsigned x0 = (unsigned)(((unsigned long long)b * n_xcorr) / tot);
  const unsigned x1 = (unsigned)(((unsigned long long)(b + 1) * n_xcorr) / tot);
  if (x1 > x0) {
    xcorr_warp_body<LDC, true>(x0, fsm, xa.proc_x, xa.proc_y, p, xa.Lw, xa.npairs, xa.corr_lag, xa.corr_val, xa.num_lags, xa.lag_idx,
                               xa.lag_sec, xa.lag_corr);
    return;
  }
  welch_warp_body(b - x0, reinterpret_cast<double*>(fsm), wa.proc_x, wa.proc_y, p, wa.max_bins, wa.nsig, 0, wa.spec_f, wa.spec_mag,
                  wa.num_bins, wa.peak_idx, wa.peak_freq, wa.peak_mag);
}

}  // namespace bpv

extern "C" int bpv_window_welch_xcorr(const double* proc_x, const double* proc_y, const bpv_window_params* p, int32_t max_bins,
                                      float* spec_f, float* spec_mag, int32_t* num_bins, int32_t* peak_idx, double* peak_freq,
                                      double* peak_mag, float* corr_lag, float* corr_val, int32_t* num_lags, int32_t* lag_idx,
                                      double* lag_sec, double* lag_corr, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(p && proc_x && proc_y && num_bins && peak_idx && peak_freq && peak_mag && num_lags && lag_idx && lag_sec && lag_corr,
              BPV_E_INVALID, "bpv_window_welch_xcorr: NULL pointer");
  BPV_REQUIRE((spec_f == nullptr) == (spec_mag == nullptr) && (corr_lag == nullptr) == (corr_val == nullptr), BPV_E_INVALID,
              "bpv_window_welch_xcorr: spec_f/spec_mag and corr_lag/corr_val must be set or NULL in pairs");
  BPV_REQUIRE(p->transform == BPV_PGRAM_WELCH, BPV_E_UNSUPPORTED, "bpv_window_welch_xcorr: the fused launch is for PGRAM_WELCH (transform %d)",
              p->transform);
  const int W = p->window, P = p->R * (p->R - 1) / 2;
  const long long nsig = (long long)p->S * p->jobs_per_stream * p->R, npairs = (long long)p->S * p->jobs_per_stream * P;
  BPV_REQUIRE(W > 0 && nsig > 0 && max_bins > 0, BPV_E_INVALID, "bpv_window_welch_xcorr: bad sizes");
  const XwLayout Lw = xw_layout(W);
  const long long n_welch = (nsig + WELCH_WPB - 1) / WELCH_WPB, n_xcorr = (npairs + 3) / 4;
  // one grid needs CTAs of one shape: four warps of either role, coarse values in registers (windows up to 320 samples);
  // anything else is the two stand-alone launches, one after the other on the caller's stream
  const bool fused = P > 0 && W <= 320 && Lw.one_round && 4 * Lw.total <= 200 * 1024 && n_welch + n_xcorr < (1LL << 31);
  if (!fused) {
    if (int rc = bpv_window_spectrum(proc_x, proc_y, p, max_bins, nullptr, 0, spec_f, spec_mag, num_bins, peak_idx, peak_freq, peak_mag, stream))
      return rc;
    return bpv_window_xcorr(proc_x, proc_y, p, corr_lag, corr_val, num_lags, lag_idx, lag_sec, lag_corr, stream);
  }
  const int need = (W < 256 ? W : 256) / 2 + 1;
  BPV_REQUIRE(!spec_mag || max_bins >= need, BPV_E_INVALID, "bpv_window_welch_xcorr: max_bins %d < %d", max_bins, need);
  const size_t sm_w = (size_t)(512 + WELCH_WPB * welch_warp_doubles(W)) * sizeof(double), sm_x = (size_t)4 * Lw.total;
  const size_t smem = sm_w > sm_x ? sm_w : sm_x;
  const unsigned grid = (unsigned)(n_welch + n_xcorr);
  const WelchArgs wa{proc_x, proc_y, max_bins, nsig, spec_f, spec_mag, num_bins, peak_idx, peak_freq, peak_mag};
  const XcorrArgs xa{proc_x, proc_y, Lw, npairs, corr_lag, corr_val, num_lags, lag_idx, lag_sec, lag_corr};
#define BPV_WX(LDC)                                                                                                          \
  do {                                                                                                                       \
    if (int rc = ensure_dyn_smem((const void*)welch_xcorr_kernel<LDC>, smem)) return rc;                                     \
    welch_xcorr_kernel<LDC><<<grid, 128, smem, (cudaStream_t)stream>>>(*p, wa, xa, (unsigned)n_welch, (unsigned)n_xcorr);    \
  } while (0)
  if (Lw.LD == 63) BPV_WX(63);
  else if (Lw.LD == 53) BPV_WX(53);
  else BPV_WX(0);
#undef BPV_WX
  return check_launch("bpv_window_welch_xcorr");
}
