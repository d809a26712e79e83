// F1 — ROI sampling: uint8 HWC-BGR frames -> exact per-ROI (sumB, sumG, sumR, N) and the float64
// sample the reference's np.mean produces (signal_processor.py:176-189).
//
// HBM-bound byte work (no tensor cores).  One thread group (32 / 128 / 256 threads) per ROI.  The
// ROI is walked as a flat list of 16-byte-aligned vectors (rows x vectors-per-row); every thread
// keeps UNROLL independent 128-bit streaming loads in flight.  A vector holds 16 interleaved BGR
// bytes whose channel phase depends on its offset from the row's first pixel; the three channel
// sums come from 12 dp4a against constant 0/1 byte selectors and are rotated by the phase.
// Head/tail vectors are byte-masked.  Integer partials are reduced with warp shuffles (uint64).
#include "common.cuh"

namespace bpv {

struct RoiArgs {
  const uint8_t* frames;
  const uint8_t* const* frame_ptrs;
  long long frame_stride, row_stride;
  int H, W, R, mode;
  long long num_rois;
  const int32_t* boxes;
  unsigned long long* out_sums;
  double* out_value;
};

// Python seq[a:b] normalisation on an axis of length L -> [s, e), e >= s.
__device__ __forceinline__ void py_slice(int a, int b, int L, int& s, int& e) {
  long long aa = a, bb = b;
  if (aa < 0) { aa += L; if (aa < 0) aa = 0; } else if (aa > L) aa = L;
  if (bb < 0) { bb += L; if (bb < 0) bb = 0; } else if (bb > L) bb = L;
  if (bb < aa) bb = aa;
  s = (int)aa; e = (int)bb;
}

__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// bytes [l, h) of a 32-bit word (l, h may lie outside 0..4)
__device__ __forceinline__ uint32_t byte_mask(int l, int h) {
  uint32_t lo = l <= 0 ? 0xffffffffu : (l >= 4 ? 0u : 0xffffffffu << (8 * l));
  uint32_t hi = h >= 4 ? 0xffffffffu : (h <= 0 ? 0u : 0xffffffffu >> (8 * (4 - h)));
  return lo & hi;
}

// selector q keeps byte b of the vector iff (q + b) % 3 == 0
__device__ __forceinline__ void channel_sums(const uint4& v, uint32_t& s0, uint32_t& s1, uint32_t& s2) {
  s0 = __dp4a(v.x, 0x01000001u, __dp4a(v.y, 0x00010000u, __dp4a(v.z, 0x00000100u, __dp4a(v.w, 0x01000001u, 0u))));
  s1 = __dp4a(v.x, 0x00010000u, __dp4a(v.y, 0x00000100u, __dp4a(v.z, 0x01000001u, __dp4a(v.w, 0x00010000u, 0u))));
  s2 = __dp4a(v.x, 0x00000100u, __dp4a(v.y, 0x01000001u, __dp4a(v.z, 0x00010000u, __dp4a(v.w, 0x00000100u, 0u))));
}

template <int GROUP, int UNROLL>
__global__ void __launch_bounds__(256) roi_sample_kernel(const RoiArgs a) {
  constexpr int GROUPS_PER_BLOCK = 256 / GROUP;
  constexpr int WARPS_PER_GROUP = GROUP / 32;
  const int grp = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
  const long long roi = (long long)blockIdx.x * GROUPS_PER_BLOCK + grp;
  const bool live = roi < a.num_rois;

  int xs = 0, xe = 0, ys = 0, ye = 0;
  bool has_box = false;
  const uint8_t* base = nullptr;
  if (live) {
    const int4 b = __ldg(reinterpret_cast<const int4*>(a.boxes) + roi);
    has_box = b.x != BPV_NO_BOX;
    if (has_box) {
      py_slice(b.x, b.z, a.W, xs, xe);
      py_slice(b.y, b.w, a.H, ys, ye);
      const long long f = roi / a.R;
      const uint8_t* fp = a.frame_ptrs ? a.frame_ptrs[f] : a.frames + f * a.frame_stride;
      base = fp + (long long)ys * a.row_stride + (long long)xs * 3;
    }
  }
  const int nrows = ye - ys, row_bytes = (xe - xs) * 3;
  uint32_t sB = 0, sG = 0, sR = 0;
  if (nrows > 0 && row_bytes > 0) {
    const uint32_t vpr = (uint32_t)(row_bytes + 14) / 16u + 1u;  // most aligned vectors a row can touch
    const uint32_t total = (uint32_t)nrows * vpr;
    uint32_t row = (uint32_t)gt / vpr, v = (uint32_t)gt % vpr;
    const uint32_t dq = (uint32_t)GROUP / vpr, dr = (uint32_t)GROUP % vpr;
    for (uint32_t idx = gt; idx < total; idx += GROUP * UNROLL) {
      uint4 d[UNROLL];
      int lo[UNROLL], hi[UNROLL], ph[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        hi[u] = 0;
        if (idx + u * GROUP < total) {
          const uintptr_t a0 = reinterpret_cast<uintptr_t>(base) + (uintptr_t)row * (uintptr_t)a.row_stride;
          const int off = (int)(a0 & 15);
          const int rel = 16 * (int)v - off;  // vector byte 0 relative to the row's first pixel byte
          const int e = row_bytes - rel;      // bytes of the row at/after vector byte 0
          if (e > 0) {
            lo[u] = rel < 0 ? -rel : 0;
            hi[u] = e < 16 ? e : 16;
            ph[u] = (rel + 15) % 3;
            d[u] = ld_stream_v4(reinterpret_cast<const void*>(a0 + rel));
          }
        }
        row += dq; v += dr;
        if (v >= vpr) { v -= vpr; ++row; }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (hi[u] > 0) {
          uint4 w = d[u];
          if (lo[u] > 0 || hi[u] < 16) {
            w.x &= byte_mask(lo[u], hi[u]);
            w.y &= byte_mask(lo[u] - 4, hi[u] - 4);
            w.z &= byte_mask(lo[u] - 8, hi[u] - 8);
            w.w &= byte_mask(lo[u] - 12, hi[u] - 12);
          }
          uint32_t s0, s1, s2;
          channel_sums(w, s0, s1, s2);
          const int p = ph[u];
          sB += p == 0 ? s0 : (p == 1 ? s1 : s2);
          sG += p == 0 ? s2 : (p == 1 ? s0 : s1);
          sR += p == 0 ? s1 : (p == 1 ? s2 : s0);
        }
      }
    }
  }

  unsigned long long tB = warp_sum_u64(sB), tG = warp_sum_u64(sG), tR = warp_sum_u64(sR);
  if (WARPS_PER_GROUP > 1) {
    __shared__ unsigned long long part[8][3];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { part[wid][0] = tB; part[wid][1] = tG; part[wid][2] = tR; }
    __syncthreads();
    if (gt == 0) {
      tB = tG = tR = 0;
      const int w0 = grp * WARPS_PER_GROUP;
#pragma unroll
      for (int w = 0; w < WARPS_PER_GROUP; ++w) { tB += part[w0 + w][0]; tG += part[w0 + w][1]; tR += part[w0 + w][2]; }
    }
  }
  if (live && gt == 0) {
    const unsigned long long N = (unsigned long long)(nrows > 0 ? nrows : 0) * (unsigned long long)(xe - xs);
    if (a.out_sums) {
      ulonglong4 o; o.x = tB; o.y = tG; o.z = tR; o.w = N;
      *reinterpret_cast<ulonglong4*>(a.out_sums + 4 * roi) = o;
    }
    double val;
    if (!has_box || N == 0) val = nan_f64();
    else if (a.mode == BPV_GREEN) val = (double)tG / (double)N;
    else val = (double)(2 * (long long)tG - (long long)tB - (long long)tR + 2 * (long long)N) / (double)(4 * N);
    a.out_value[roi] = val;
  }
}

}  // namespace bpv

extern "C" int bpv_roi_sample_u8(const uint8_t* frames, const uint8_t* const* frame_ptrs,
                                 int64_t frame_stride_bytes, int64_t row_stride_bytes,
                                 int32_t H, int32_t W, int64_t num_frames,
                                 const int32_t* boxes, int32_t R, int32_t mode,
                                 uint64_t* out_sums, double* out_value,
                                 int64_t roi_pixels_hint, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(frames || frame_ptrs, BPV_E_INVALID, "bpv_roi_sample_u8: frames and frame_ptrs are both NULL");
  BPV_REQUIRE(boxes && out_value, BPV_E_INVALID, "bpv_roi_sample_u8: NULL boxes/out_value");
  BPV_REQUIRE(H > 0 && W > 0 && R > 0 && num_frames >= 0, BPV_E_INVALID, "bpv_roi_sample_u8: bad H/W/R/num_frames");
  BPV_REQUIRE((int64_t)H * W * 3 < (1ll << 31), BPV_E_TOO_LARGE, "bpv_roi_sample_u8: frame larger than 2 GiB");
  BPV_REQUIRE(row_stride_bytes >= 3ll * W, BPV_E_INVALID, "bpv_roi_sample_u8: row_stride_bytes < 3*W");
  BPV_REQUIRE(mode == BPV_GREEN || mode == BPV_CHROM_GREEN, BPV_E_UNSUPPORTED,
              "bpv_roi_sample_u8: unknown color channel %d (NotImplementedError, signal_processor.py:185)", mode);
  if (num_frames == 0) return 0;
  RoiArgs a{frames, frame_ptrs, frame_stride_bytes, row_stride_bytes, H, W, R, mode,
            num_frames * R, boxes, (unsigned long long*)out_sums, out_value};
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = a.num_rois;
  // threads per ROI: ~>= 4 vectors per thread before widening the group
  const long long px = roi_pixels_hint > 0 ? roi_pixels_hint : 4096;
  if (px * 3 <= 32 * 16 * 8) {
    roi_sample_kernel<32, 4><<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a);
  } else if (px * 3 <= 128 * 16 * 16) {
    roi_sample_kernel<128, 4><<<(unsigned)((n + 1) / 2), 256, 0, st>>>(a);
  } else {
    roi_sample_kernel<256, 4><<<(unsigned)n, 256, 0, st>>>(a);
  }
  return check_launch("bpv_roi_sample_u8");
}
