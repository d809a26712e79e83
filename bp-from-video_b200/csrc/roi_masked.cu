// F1 through a segmentation mask — SURVEY.md 8(f) row 4.  The reference computes a per-pixel category mask for
// every frame (InferenceRunner.run_person_segmenter: category_mask of the selfie-multiclass segmenter, uint8 [H, W],
// inference_runner.py:154-166) but only uses it for drawing (drawer.py:95-99); sampling a ROI only over the pixels
// of one category (3 = face skin for the forehead box, 2 = body skin for the palm box) is the extension the survey
// names.  Semantics, as a numpy statement (the CPU checker under tests/ restates exactly this):
//     sel = mask[y0:y1, x0:x1] == category[r];   sample = np.mean(channel(frame[y0:y1, x0:x1])[sel])
// with the same Python slice semantics as sample_signal (signal_processor.py:176-189); no selected pixel -> NaN.
// The result is the exact integer sums (sumB, sumG, sumR, N) over the selected pixels and the float64 np.mean makes
// of them.
//
// One 128-thread CTA per ROI; a thread owns 4-pixel groups aligned to 4 pixels: ONE 32-bit mask word and THREE 32-bit
// frame words (12 BGR bytes) per group, channel sums by dp4a against selectors built from the four category tests.
// Frames / masks whose base or strides are not 4-byte aligned take the byte-load path.  HBM bound: 4 B per ROI pixel.
#include "common.cuh"

namespace bpv {

__device__ __forceinline__ void py_slice_m(int a, int b, int L, int& s, int& e) {
  long long aa = a, bb = b;
  if (aa < 0) { aa += L; if (aa < 0) aa = 0; } else if (aa > L) aa = L;
  if (bb < 0) { bb += L; if (bb < 0) bb = 0; } else if (bb > L) bb = L;
  if (bb < aa) bb = aa;
  s = (int)aa; e = (int)bb;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(128) roi_masked_kernel(const uint8_t* __restrict__ frames, long long frame_stride,
                                                         long long row_stride, const uint8_t* __restrict__ masks,
                                                         long long mask_frame_stride, long long mask_row_stride,
                                                         int H, int W, int R, int mode, const int32_t* __restrict__ boxes,
                                                         const int32_t* __restrict__ categories,
                                                         unsigned long long* __restrict__ out_sums, double* __restrict__ out_value) {
  constexpr int THREADS = 128, WARPS = THREADS / 32;
  const int gt = threadIdx.x;
  const long long roi = blockIdx.x;
  int xs = 0, xe = 0, ys = 0, ye = 0;
  const int4 b = __ldg(reinterpret_cast<const int4*>(boxes) + roi);
  const bool has_box = b.x != BPV_NO_BOX;
  if (has_box) { py_slice_m(b.x, b.z, W, xs, xe); py_slice_m(b.y, b.w, H, ys, ye); }
  const unsigned cat = (unsigned)__ldg(categories + (int)(roi % R)) & 0xffu;
  const uint8_t* fp = frames + (roi / R) * frame_stride;
  const uint8_t* mp = masks + (roi / R) * mask_frame_stride;
  const int nrows = ye - ys, ncols = xe - xs;
  uint32_t sB = 0, sG = 0, sT = 0, cnt = 0;     // blue, green, all three channels, selected pixels
  if (nrows > 0 && ncols > 0) {
    const int x4 = xs & ~3;
    const int gpr = (xe - x4 + 3) >> 2;         // 4-pixel groups per row
    int rps, r0, g0;
    if (gpr >= THREADS) { rps = 1; r0 = 0; g0 = gt; }
    else { rps = THREADS / gpr; r0 = gt / gpr; g0 = gt - r0 * gpr; if (r0 >= rps) g0 = gpr; }
    for (int g = g0; g < gpr; g += THREADS) {
      const int x = x4 + 4 * g;
      // pixels of the group inside [xs, xe): loop invariant for the thread
      unsigned inx = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) inx |= (x + e >= xs && x + e < xe) ? 1u << e : 0u;
      for (int r = r0; r < nrows; r += rps) {
        const int y = ys + r;
        const uint8_t* mrow = mp + (long long)y * mask_row_stride + x;
        const uint8_t* frow = fp + (long long)y * row_stride + 3LL * x;
        uint32_t mw, w0, w1, w2;
        if (ALIGNED) {
          mw = __ldg(reinterpret_cast<const uint32_t*>(mrow));
          w0 = __ldg(reinterpret_cast<const uint32_t*>(frow));
          w1 = __ldg(reinterpret_cast<const uint32_t*>(frow) + 1);
          w2 = __ldg(reinterpret_cast<const uint32_t*>(frow) + 2);
        } else {
          // byte loads; bytes of pixels outside the frame's row are never touched (x + e < W for every pixel in range)
          mw = w0 = w1 = w2 = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (inx >> e & 1) {
              mw |= (uint32_t)mrow[e] << (8 * e);
              const uint32_t px = (uint32_t)frow[3 * e] | (uint32_t)frow[3 * e + 1] << 8 | (uint32_t)frow[3 * e + 2] << 16;
              // pixel e occupies bytes 3e .. 3e+2 of the 12-byte group
              if (e == 0) w0 |= px;
              else if (e == 1) { w0 |= px << 24; w1 |= px >> 8; }
              else if (e == 2) { w1 |= px << 16; w2 |= px >> 16; }
              else w2 |= px << 8;
            } else {
              mw |= (cat ^ 0xffu) << (8 * e);    // any value != category
            }
          }
        }
        // f_e = 1 iff pixel e is in range and its mask byte equals the category
        const unsigned f0 = (inx & 1u) && ((mw & 0xffu) == cat);
        const unsigned f1 = (inx >> 1 & 1u) && ((mw >> 8 & 0xffu) == cat);
        const unsigned f2 = (inx >> 2 & 1u) && ((mw >> 16 & 0xffu) == cat);
        const unsigned f3 = (inx >> 3 & 1u) && ((mw >> 24) == cat);
        // 12 bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        const uint32_t t0 = f0 * 0x00010101u | f1 << 24, t1 = f1 * 0x00000101u | f2 * 0x01010000u, t2 = f2 | f3 * 0x01010100u;
        const uint32_t g0s = f0 << 8, g1s = f1 | f2 << 24, g2s = f3 << 16;
        const uint32_t b0s = f0 | f1 << 24, b1s = f2 << 16, b2s = f3 << 8;
        sT = __dp4a(w0, t0, __dp4a(w1, t1, __dp4a(w2, t2, sT)));
        sG = __dp4a(w0, g0s, __dp4a(w1, g1s, __dp4a(w2, g2s, sG)));
        sB = __dp4a(w0, b0s, __dp4a(w1, b1s, __dp4a(w2, b2s, sB)));
        cnt += f0 + f1 + f2 + f3;
      }
    }
  }
  unsigned long long tB = warp_sum_u64(sB), tG = warp_sum_u64(sG), tT = warp_sum_u64(sT), tN = warp_sum_u64(cnt);
  __shared__ unsigned long long part[WARPS][4];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { part[wid][0] = tB; part[wid][1] = tG; part[wid][2] = tT; part[wid][3] = tN; }
  __syncthreads();
  if (gt == 0) {
    tB = tG = tT = tN = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { tB += part[w][0]; tG += part[w][1]; tT += part[w][2]; tN += part[w][3]; }
    if (out_sums) {
      ulonglong4 o; o.x = tB; o.y = tG; o.z = tT - tG - tB; o.w = tN;
      *reinterpret_cast<ulonglong4*>(out_sums + 4 * roi) = o;
    }
    double val;
    if (!has_box || tN == 0) val = nan_f64();
    else if (mode == BPV_GREEN) val = (double)tG / (double)tN;
    else val = (double)(3 * (long long)tG - (long long)tT + 2 * (long long)tN) / (double)(4 * tN);   // 2G - B - R + 2N
    out_value[roi] = val;
  }
}

}  // namespace bpv

extern "C" int bpv_roi_sample_masked_u8(const uint8_t* frames, int64_t frame_stride_bytes, int64_t row_stride_bytes,
                                        const uint8_t* masks, int64_t mask_frame_stride_bytes, int64_t mask_row_stride_bytes,
                                        int32_t H, int32_t W, int64_t num_frames, const int32_t* boxes, int32_t R,
                                        const int32_t* categories, int32_t mode, uint64_t* out_sums, double* out_value,
                                        void* stream) {
  using namespace bpv;
  BPV_REQUIRE(frames && masks && boxes && categories && out_value, BPV_E_INVALID, "bpv_roi_sample_masked_u8: NULL pointer");
  BPV_REQUIRE(H > 0 && W > 0 && R > 0 && num_frames >= 0, BPV_E_INVALID, "bpv_roi_sample_masked_u8: bad sizes");
  BPV_REQUIRE(row_stride_bytes >= 3LL * W && mask_row_stride_bytes >= W, BPV_E_INVALID,
              "bpv_roi_sample_masked_u8: row stride < 3*W or mask row stride < W");
  BPV_REQUIRE(mode == BPV_GREEN || mode == BPV_CHROM_GREEN, BPV_E_UNSUPPORTED,
              "bpv_roi_sample_masked_u8: unknown color channel %d (NotImplementedError, signal_processor.py:185)", mode);
  BPV_REQUIRE(num_frames * (int64_t)R <= INT32_MAX, BPV_E_TOO_LARGE, "bpv_roi_sample_masked_u8: more than 2^31-1 ROIs in one call");
  if (num_frames == 0) return 0;
  const long long n = num_frames * R;
  cudaStream_t st = (cudaStream_t)stream;
  // the 32-bit path reads whole words around the ROI: every word must lie inside the row, i.e. strides padded to 4
  const bool aligned = (((uintptr_t)frames | (uintptr_t)frame_stride_bytes | (uintptr_t)row_stride_bytes |
                         (uintptr_t)masks | (uintptr_t)mask_frame_stride_bytes | (uintptr_t)mask_row_stride_bytes) & 3) == 0 &&
                       mask_row_stride_bytes >= ((W + 3) & ~3) && row_stride_bytes >= 3LL * ((W + 3) & ~3);
  if (aligned)
    roi_masked_kernel<true><<<(unsigned)n, 128, 0, st>>>(frames, frame_stride_bytes, row_stride_bytes, masks, mask_frame_stride_bytes,
                                                         mask_row_stride_bytes, H, W, R, mode, boxes, categories,
                                                         (unsigned long long*)out_sums, out_value);
  else
    roi_masked_kernel<false><<<(unsigned)n, 128, 0, st>>>(frames, frame_stride_bytes, row_stride_bytes, masks, mask_frame_stride_bytes,
                                                          mask_row_stride_bytes, H, W, R, mode, boxes, categories,
                                                          (unsigned long long*)out_sums, out_value);
  return check_launch("bpv_roi_sample_masked_u8");
}
