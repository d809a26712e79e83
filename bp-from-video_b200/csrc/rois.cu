// ROI geometry on the device for batched landmark tensors — SURVEY.md §8(f) row 1, the step in front of F1:
//   SignalProcessor.calc_rois              (signal_processor.py:133-155)
//   sg_roi.add_samples + get_means(as_int) (signal_processor.py:304-305; signal_data.py:31-35, 60-63)
// One thread per (stream, ROI) walks the T frames of the step in order (the smoothing history is sequential).
// All arithmetic is the reference's float64 arithmetic: mean of the selected landmark points, numpy round
// (half to even), anchor + relative_bbox * bbox size with separate multiply and add (no FMA), Python round
// (half to even), nanmean over the history, round again.
#include "common.cuh"

namespace bpv {

constexpr int MAX_ROI_POINTS = 8;

__global__ void calc_rois_kernel(const uint8_t* __restrict__ present, const int32_t* __restrict__ bbox,
                                 const int32_t* __restrict__ points, const int32_t* __restrict__ num_points,
                                 const double* __restrict__ rel_bbox, int S, int T, int R, int K, int H, long long g0,
                                 double* __restrict__ hist, double* __restrict__ locations, double* __restrict__ smoothed,
                                 int32_t* __restrict__ boxes) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * R) return;
  const int s = idx / R, r = idx % R;
  const int np = num_points[r];
  const double l = rel_bbox[4 * r], tp = rel_bbox[4 * r + 1], rt = rel_bbox[4 * r + 2], bt = rel_bbox[4 * r + 3];
  double* hs = hist + (long long)idx * H * 6;
  for (int t = 0; t < T; ++t) {
    const long long e = ((long long)s * T + t) * R + r;
    double loc[6];
    if (present[e]) {
      const int32_t* bb = bbox + e * 4;
      const int32_t* pt = points + e * K * 2;
      double sx = 0.0, sy = 0.0;
      for (int k = 0; k < np; ++k) { sx += (double)pt[2 * k]; sy += (double)pt[2 * k + 1]; }
      const double x = rint(sx / (double)np), y = rint(sy / (double)np);        // np.mean(...).round()
      const double bw = (double)(bb[2] - bb[0]), bh = (double)(bb[3] - bb[1]);
      loc[0] = x; loc[1] = y;
      loc[2] = rint(__dadd_rn(x, __dmul_rn(l, bw)));                              // int(round(x + left * bbox width))
      loc[3] = rint(__dadd_rn(y, __dmul_rn(tp, bh)));
      loc[4] = rint(__dadd_rn(x, __dmul_rn(rt, bw)));
      loc[5] = rint(__dadd_rn(y, __dmul_rn(bt, bh)));
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) loc[c] = nan_f64();                              // (nan,) * 6
    }
    double* slot = hs + (int)((g0 + t) % H) * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) slot[c] = loc[c];
    if (locations)
      for (int c = 0; c < 6; ++c) locations[e * 6 + c] = loc[c];
    // Signal.get_mean(as_int=True): nanmean over the rows whose 6 entries are all finite, round half to even
    double sum[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int h = 0; h < H; ++h) {
      const double* row = hs + h * 6;
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 6; ++c) ok = ok && isfinite(row[c]);
      if (ok) {
        ++cnt;
#pragma unroll
        for (int c = 0; c < 6; ++c) sum[c] += row[c];
      }
    }
    double m[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) m[c] = cnt ? rint(sum[c] / (double)cnt) : nan_f64();
    if (smoothed)
      for (int c = 0; c < 6; ++c) smoothed[e * 6 + c] = m[c];
    int32_t* bo = boxes + e * 4;
    if (cnt) { bo[0] = (int32_t)m[2]; bo[1] = (int32_t)m[3]; bo[2] = (int32_t)m[4]; bo[3] = (int32_t)m[5]; }
    else { bo[0] = BPV_NO_BOX; bo[1] = bo[2] = bo[3] = 0; }
  }
}

}  // namespace bpv

extern "C" int bpv_calc_rois(const uint8_t* present, const int32_t* bbox, const int32_t* points,
                             const int32_t* num_points, const double* rel_bbox,
                             int32_t S, int32_t T, int32_t R, int32_t K, int32_t H, int64_t g0,
                             double* hist, double* locations, double* smoothed, int32_t* boxes, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(present && bbox && points && num_points && rel_bbox && hist && boxes, BPV_E_INVALID, "bpv_calc_rois: NULL pointer");
  BPV_REQUIRE(S > 0 && T > 0 && R > 0 && K > 0 && K <= MAX_ROI_POINTS && H > 0 && g0 >= 0, BPV_E_INVALID, "bpv_calc_rois: bad sizes");
  const int n = S * R;
  calc_rois_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(present, bbox, points, num_points, rel_bbox, S, T, R, K, H, g0,
                                                                     hist, locations, smoothed, boxes);
  return check_launch("bpv_calc_rois");
}

// ---------------------------------------------------------------------------------------------
// VideoReader view in front of F1 — SURVEY.md §8(f) row 2 (video_reader.py:97-103): the reference crops the decoded
// frame to a centred portrait window (frame[:, left:right]) and mirrors it (cv2.flip(frame, 1)) BEFORE the signal
// path sees it; ROI boxes are expressed in that view.  A ROI mean is invariant under mirroring the ROI, so instead
// of materialising the cropped / flipped frame the box is mapped back onto the decoded frame in HBM:
//   view column c  <->  source column  left + c            (no flip)
//                       left + (view_w - 1 - c)            (flip)
// The box is first normalised with Python slice semantics against the VIEW width (negative indices wrap around the
// view, stops clamp), so the mapped box is a plain in-range box of the source frame; rows are untouched.
namespace bpv {
__global__ void view_boxes_kernel(const int32_t* __restrict__ boxes, long long n, int view_w, int view_h, int left, int flip,
                                  int32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 b = reinterpret_cast<const int4*>(boxes)[i];
  int4 o = b;
  if (b.x != BPV_NO_BOX) {
    long long xs = b.x, xe = b.z, ys = b.y, ye = b.w;
    if (xs < 0) { xs += view_w; if (xs < 0) xs = 0; } else if (xs > view_w) xs = view_w;
    if (xe < 0) { xe += view_w; if (xe < 0) xe = 0; } else if (xe > view_w) xe = view_w;
    if (xe < xs) xe = xs;
    if (ys < 0) { ys += view_h; if (ys < 0) ys = 0; } else if (ys > view_h) ys = view_h;
    if (ye < 0) { ye += view_h; if (ye < 0) ye = 0; } else if (ye > view_h) ye = view_h;
    if (ye < ys) ye = ys;
    if (flip) { const long long t = view_w - xe; xe = view_w - xs; xs = t; }
    o.x = (int)(xs + left); o.z = (int)(xe + left); o.y = (int)ys; o.w = (int)ye;
  }
  reinterpret_cast<int4*>(out)[i] = o;
}
}  // namespace bpv

extern "C" int bpv_view_boxes(const int32_t* boxes, int64_t num_boxes, int32_t view_w, int32_t view_h, int32_t left,
                              int32_t flip_horizontally, int32_t* out, void* stream) {
  using namespace bpv;
  BPV_REQUIRE(boxes && out, BPV_E_INVALID, "bpv_view_boxes: NULL pointer");
  BPV_REQUIRE(num_boxes >= 0 && view_w > 0 && view_h > 0 && left >= 0, BPV_E_INVALID, "bpv_view_boxes: bad sizes");
  if (num_boxes == 0) return 0;
  view_boxes_kernel<<<(unsigned)((num_boxes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes, num_boxes, view_w, view_h, left,
                                                                                         flip_horizontally, out);
  return check_launch("bpv_view_boxes");
}
