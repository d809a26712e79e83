"""CPU oracle for the bp-from-video signal path.  TEST INFRASTRUCTURE ONLY.

This is a numpy/scipy restatement of the reference's hot path
(`/root/reference/signal_processor.py`, `signal_data.py`, `roi.py`) written as plain
functions over arrays.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product (`bp-from-video_b200/`)
never does.

Where the arithmetic lives
--------------------------
The reference is 100 % Python; every number on the path is produced by numpy / scipy
calls (requirements.txt:1-5, unpinned).  The effective pins are the versions in this image:
scipy 1.18.1, numpy 2.3.5.  This oracle makes the *same* third-party calls with the *same*
kwargs in the *same* order as the reference call sites cited on every function, so it is
bit-identical to the reference on the same inputs.

Parity pin
----------
The reference has no tests or golden vectors of its own ("parity unpinned by the
reference").  The oracle is pinned instead against outputs of the reference itself, run in the
build container by `tests/golden/make_golden.py` (imports `/root/reference` unmodified) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` requires exact equality.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import scipy.fft
import scipy.interpolate
import scipy.signal

# ---------------------------------------------------------------------------------------------
# method / transform / channel codes (shared numbering with include/bpv.h)
# ---------------------------------------------------------------------------------------------
GREEN, CHROM_GREEN = 0, 1
DIFF_1, DIFF_2, INTERP_LINEAR, INTERP_CUBIC, DETREND_CONST, DETREND_LINEAR, FILTER_BUTTER, FILTER_FIR = range(1, 9)
DFT_RFFT, PGRAM_WELCH, PGRAM_LS = 1, 2, 3

DEFAULTS = dict(  # signal_processor.py:45-72
    butter_order=16, butter_min_bw=0.1, fir_taps=127, fir_df=0.3,
    min_freq=0.8, max_freq=4.0, ls_num_freqs=None,
)


def _params(kw):
    p = dict(DEFAULTS)
    p.update(kw)
    return p


# ---------------------------------------------------------------------------------------------
# ROI geometry and sampling
# ---------------------------------------------------------------------------------------------
def calc_roi(detections, landmark_indices, relative_bbox):
    """signal_processor.py:142-153 — anchor = rounded mean of the selected landmarks of the
    largest detection; box corners = anchor + relative_bbox * bbox size, Python round()."""
    if len(detections) == 0:
        return (np.nan,) * 6
    bbox, points = detections[0]
    pp = np.squeeze(np.mean([points[i] for i in landmark_indices], axis=0))
    x, y = pp.round().astype(int)
    l, t, r, b = relative_bbox
    x0 = int(round(x + l * (bbox[2] - bbox[0])))
    y0 = int(round(y + t * (bbox[3] - bbox[1])))
    x1 = int(round(x + r * (bbox[2] - bbox[0])))
    y1 = int(round(y + b * (bbox[3] - bbox[1])))
    return (x, y, x0, y0, x1, y1)


def smooth_roi(history):
    """signal_data.py:60-63 with as_int=True — nanmean over the last boxes, half-to-even
    round; all-NaN history returns the last (NaN) row."""
    y = np.asarray(history, dtype=float).reshape(-1, 6)
    w = np.isfinite(y).all(axis=1)
    if not w.any():
        return y[-1]
    return np.squeeze(np.nanmean(y, axis=0)).round().astype(int)


def py_slice(a, b, n):
    """Python `seq[a:b]` index normalisation for a length-n axis -> (start, stop), stop>=start."""
    a, b, _ = slice(int(a), int(b)).indices(int(n))
    return a, max(a, b)


def roi_sums(frame, box):
    """Exact integer restatement of signal_processor.py:176-186: (sumB, sumG, sumR, N) over
    frame[y0:y1, x0:x1, :] with Python slice semantics."""
    x0, y0, x1, y1 = box
    H, W = frame.shape[:2]
    ra, rb = py_slice(y0, y1, H)
    ca, cb = py_slice(x0, x1, W)
    tile = frame[ra:rb, ca:cb, :].astype(np.uint64)
    n = (rb - ra) * (cb - ca)
    s = tile.reshape(-1, 3).sum(axis=0) if n else np.zeros(3, np.uint64)
    return int(s[0]), int(s[1]), int(s[2]), int(n)


def resize_linear_u8(src, dsize_w, dsize_h):
    """cv2.resize(src, (dsize_w, dsize_h)) with the default INTER_LINEAR for uint8 HWC images — what VideoReader applies to
    file input when target_res is set (video_reader.py:95-96) — restated in integer arithmetic
    (opencv/modules/imgproc/src/resize.cpp: resizeGeneric_ / HResizeLinear / VResizeLinear with 11-bit coefficients;
    exact 2x decimation takes the INTER_AREA fast path).  Pinned against cv2 in tests/test_oracle_golden.py."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    if sw == 2 * dsize_w and sh == 2 * dsize_h:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def frac(dn, sn):
        scale = 1.0 / (float(dn) / float(sn))                       # scale = 1 / inv_scale, as cv::resize computes it
        f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
        s0 = np.floor(f).astype(np.int64)
        return s0, (f - s0.astype(np.float32)).astype(np.float32)

    def icoef(fr):                                                   # saturate_cast<short>(c * 2048) = round half to even
        return (np.rint((np.float32(1.0) - fr) * np.float32(2048)).astype(np.int64),
                np.rint(fr * np.float32(2048)).astype(np.int64))
    x0, fx = frac(dsize_w, sw)
    lo = x0 < 0
    fx[lo] = 0; x0[lo] = 0
    hi = x0 >= sw - 1
    fx[hi] = 0; x0[hi] = sw - 1
    a0, a1 = icoef(fx)
    x1 = np.minimum(x0 + 1, sw - 1)
    y0, fy = frac(dsize_h, sh)                                       # rows: weights kept, indices clipped
    b0, b1 = icoef(fy)
    y1 = np.clip(y0 + 1, 0, sh - 1)
    y0 = np.clip(y0, 0, sh - 1)
    s = src.astype(np.int64)
    h = s[:, x0] * a0[None, :, None] + s[:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (h[y0] >> 4)) >> 16) + ((b1[:, None, None] * (h[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def nv12_to_bgr(nv12, H, W):
    """uint8 [3H/2, W] NV12 -> uint8 [H, W, 3] BGR exactly as OpenCV's cvtColor(COLOR_YUV2BGR_NV12) — the conversion a
    cv2.VideoCapture applies to decoder output before the reference sees the frame (video_reader.py:93).  Integer BT.601,
    20-bit fixed point (opencv/modules/imgproc/src/color_yuv.simd.hpp: ITUR_BT_601_C*; uvToRGBuv / yRGBuvToRGBA).
    Pinned against cv2 in tests/test_oracle_golden.py."""
    nv12 = np.asarray(nv12)
    Y = nv12[:H, :W].astype(np.int64)
    UV = nv12[H:H + H // 2, :W].reshape(H // 2, W // 2, 2).astype(np.int64)
    U = np.repeat(np.repeat(UV[..., 0], 2, axis=0), 2, axis=1) - 128
    V = np.repeat(np.repeat(UV[..., 1], 2, axis=0), 2, axis=1) - 128
    y = np.maximum(0, Y - 16) * 1220542
    half = 1 << 19
    r = (y + half + 1673527 * V) >> 20
    g = (y + half - 852492 * V - 409993 * U) >> 20
    b = (y + half + 2116026 * U) >> 20
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


def roi_sample(frame, sroi, channel=GREEN):
    """signal_processor.py:176-189, called exactly as the reference does (numpy slicing +
    np.mean in float64).  `sroi` is the 6-tuple Location; NaN anywhere -> NaN."""
    if np.isnan(sroi).any():
        return np.nan
    _, _, x0, y0, x1, y1 = sroi
    roi_bgr = frame[y0:y1, x0:x1, :]
    if channel == GREEN:
        px = roi_bgr[..., 1]
    elif channel == CHROM_GREEN:
        px = roi_bgr[..., 1] / 2 - roi_bgr[..., 0] / 4 - roi_bgr[..., 2] / 4 + 0.5
    else:
        raise NotImplementedError
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return np.mean(px)


def roi_sample_masked(frame, mask, category, sroi, channel=GREEN):
    """SURVEY.md 8(f) row 4 — ROI sampling through a segmentation mask.  The reference never samples through its
    person-segmenter category mask (inference_runner.py:154-166 produces it, drawer.py:95-99 only draws with it), so this
    function DEFINES the extension in the reference's own idiom: sample_signal (signal_processor.py:176-189) restricted
    to the ROI pixels whose category equals `category`:  np.mean(channel(roi)[mask_roi == category]).
    Returns (value, (sumB, sumG, sumR, N_selected))."""
    if np.isnan(sroi).any():
        return np.nan, (0, 0, 0, 0)
    _, _, x0, y0, x1, y1 = sroi
    roi_bgr = frame[y0:y1, x0:x1, :]
    sel = mask[y0:y1, x0:x1] == category
    if channel == GREEN:
        px = roi_bgr[..., 1]
    elif channel == CHROM_GREEN:
        px = roi_bgr[..., 1] / 2 - roi_bgr[..., 0] / 4 - roi_bgr[..., 2] / 4 + 0.5
    else:
        raise NotImplementedError
    picked = roi_bgr[sel].astype(np.uint64)                      # [N_selected, 3]
    sums = (int(picked[:, 0].sum()), int(picked[:, 1].sum()), int(picked[:, 2].sum()), int(sel.sum()))
    import warnings
    with np.errstate(all='ignore'), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return np.mean(px[sel]), sums


def value_from_sums(sB, sG, sR, n, channel):
    """the float64 the reference's np.mean produces, from exact integer sums."""
    if n == 0:
        return np.nan
    if channel == GREEN:
        return float(sG) / float(n)
    return float(2 * sG - sB - sR + 2 * n) / float(4 * n)


# ---------------------------------------------------------------------------------------------
# window helpers (signal_data.py)
# ---------------------------------------------------------------------------------------------
def window_fs(x, mask=None):
    """signal_data.py:55-58 — 1 / nanmean(diff(x[mask])) or NaN when <2 samples."""
    x = np.asarray(x, dtype=float)
    u = np.isfinite(x) if mask is None else mask
    return 1 / np.nanmean(np.diff(x[u])) if u.sum() >= 2 else np.nan


def peak(x, y):
    """signal_data.py:65-70 as `process()` really runs it: the SignalGroup wrapper resets
    range_x to (nanmin x, nanmax x) (signal_data.py:82-86,100-102,47-49), so the search
    covers every finite sample.  Returns (x_peak, y_peak, index into the compacted u)."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    if x.size == 0:
        return np.nan, np.nan, -1
    v, w = np.isfinite(x), np.isfinite(y)
    if v.sum() < 2:
        return np.nan, np.nan, -1
    lo, hi = np.nanmin(x), np.nanmax(x)
    u = (lo <= x) & (x <= hi) & w
    if u.sum() < 2:
        return np.nan, np.nan, -1
    i = int(np.argmax(y[u]))
    return x[u][i], np.max(y[u]), int(np.flatnonzero(u)[i])


# ---------------------------------------------------------------------------------------------
# filter design
# ---------------------------------------------------------------------------------------------
def make_filter(method, fs, **kw):
    """signal_processor.py:158-173."""
    p = _params(kw)
    if method == FILTER_BUTTER:
        bands = [min(p['min_freq'], fs / 2 - 2 * p['butter_min_bw']),
                 min(p['max_freq'], fs / 2 - p['butter_min_bw'])]
        return scipy.signal.butter(p['butter_order'], bands, btype='bandpass', output='sos', fs=fs)
    if method == FILTER_FIR:
        df = p['fir_df']
        bands = [0, max(p['min_freq'] - df, df), p['min_freq'], p['max_freq'],
                 min(p['max_freq'] + df, fs / 2 - df), fs / 2]
        return scipy.signal.firls(p['fir_taps'], bands, [0, 0, 1, 1, 0, 0], fs=fs)
    raise NotImplementedError


# ---------------------------------------------------------------------------------------------
# F2: window preprocessing
# ---------------------------------------------------------------------------------------------
def preprocess(x, y, methods, **kw):
    """signal_processor.py:196-241.  x, y: full windows (NaN where missing).  Returns new
    (x, y) of the same length."""
    x, y = np.array(x, dtype=float), np.array(y, dtype=float)
    block, valid = np.isfinite(x), np.isfinite(y)
    fs = window_fs(x)
    if valid.sum() >= 2 and np.isfinite(fs):
        for m in methods:
            if m == DIFF_1:
                y[valid] = np.diff(y[valid], n=1, axis=0, prepend=y[valid][0])
            elif m == DIFF_2:
                y[valid] = np.diff(y[valid], n=2, axis=0, prepend=y[valid][:2])
            elif m == INTERP_LINEAR:
                xi, ts = np.linspace(x[block][0], x[block][-1], block.sum(), retstep=True)
                yi = np.interp(xi, x[valid], y[valid])
                x[block], y[block] = xi, yi
                valid = block
                fs = 1 / ts
            elif m == INTERP_CUBIC:
                cs = scipy.interpolate.CubicSpline(x[valid], y[valid], axis=0)
                xi, ts = np.linspace(x[block][0], x[block][-1], block.sum(), retstep=True)
                yi = cs(xi)
                x[block], y[block] = xi, yi
                valid = block
                fs = 1 / ts
            elif m == DETREND_CONST:
                y[valid] = scipy.signal.detrend(y[valid], type='constant')
            elif m == DETREND_LINEAR:
                y[valid] = scipy.signal.detrend(y[valid], type='linear')
            elif m == FILTER_BUTTER:
                sos = make_filter(m, fs, **kw)
                dpl = 3 * (2 * len(sos) + 1 - min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum()))
                padlen = valid.sum() - 1 if valid.sum() <= dpl else dpl
                y[valid] = scipy.signal.sosfiltfilt(sos, y[valid], padlen=padlen)
            elif m == FILTER_FIR:
                fir = make_filter(m, fs, **kw)
                dpl = 3 * len(fir)
                padlen = valid.sum() - 1 if valid.sum() <= dpl else dpl
                y[valid] = scipy.signal.filtfilt(fir, 1.0, y[valid], padlen=padlen)
            else:
                raise NotImplementedError
    return x, y


# ---------------------------------------------------------------------------------------------
# F3: spectra
# ---------------------------------------------------------------------------------------------
def spectrum(x, y, transform, **kw):
    """signal_processor.py:248-273.  Returns (freqs, mags) float64 arrays (empty when the
    guard fails).  `ls_num_freqs` (extension, default None = reference behaviour F=n) sets the
    Lomb-Scargle grid size for BASELINE config 3."""
    p = _params(kw)
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    valid = np.isfinite(y)
    fs = window_fs(x)
    if not (valid.sum() >= 2 and np.isfinite(fs)):
        return np.zeros(0), np.zeros(0)
    import warnings
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore')
        if transform == DFT_RFFT:
            n = len(x[valid])
            freqs = scipy.fft.rfftfreq(n, 1 / fs)
            mags = 2 * np.abs(scipy.fft.rfft(y[valid], n=n)) / n
        elif transform == PGRAM_WELCH:
            freqs, mags = scipy.signal.welch(y[valid], fs)
        elif transform == PGRAM_LS:
            n = len(x[valid])
            nf = n if p['ls_num_freqs'] is None else int(p['ls_num_freqs'])
            freqs = np.linspace(p['min_freq'], p['max_freq'], nf)
            mags = scipy.signal.lombscargle(x[valid], y[valid], freqs=freqs * 2 * np.pi,
                                            floating_mean=True, normalize=True)
        else:
            raise NotImplementedError
    return np.asarray(freqs, dtype=float), np.asarray(mags, dtype=float)


# ---------------------------------------------------------------------------------------------
# F4: cross-correlation
# ---------------------------------------------------------------------------------------------
def xcorr(xa, ya, yb):
    """signal_processor.py:280-295 — full cross-correlation of the jointly valid samples,
    normalised by max(a.a, b.b, a.b); lags in seconds from signal a's own timestamps."""
    xa, ya, yb = (np.asarray(v, dtype=float) for v in (xa, ya, yb))
    valid = np.isfinite(ya) & np.isfinite(yb)
    if valid.sum() < 2:
        return np.zeros(0), np.zeros(0)
    a, b = ya[valid], yb[valid]
    with np.errstate(all='ignore'):
        corr = scipy.signal.correlate(a, b)
        corr /= np.max([np.dot(a, a), np.dot(b, b), np.dot(a, b)])
        k = scipy.signal.correlation_lags(valid.sum(), valid.sum())
        lags = (xa[valid][-1] - xa[valid][::-1])[np.abs(k)] * np.sign(k)
    return lags, corr


# ---------------------------------------------------------------------------------------------
# whole per-frame step for one stream (signal_processor.py:302-313)
# ---------------------------------------------------------------------------------------------
class OracleStream:
    """One stream's state + the 11-step `process()` (signal_processor.py:302-313), with the
    deques replaced by NaN-prefilled numpy rings (signal_data.py:14-25)."""

    def __init__(self, num_rois=2, roi_max_samples=1, signal_max_samples=250, peak_max_samples=50,
                 channel=GREEN, methods=(FILTER_BUTTER,), transform=PGRAM_LS, **kw):
        self.R = num_rois
        self.P = math.comb(num_rois, 2)
        self.channel, self.methods, self.transform, self.kw = channel, list(methods), transform, kw
        self.roi_hist = np.full((num_rois, roi_max_samples, 6), np.nan)
        self.t = np.full(signal_max_samples, np.nan)
        self.raw = np.full((num_rois, signal_max_samples), np.nan)
        self.bpm_t = np.full(peak_max_samples, np.nan)
        self.bpm = np.full((num_rois, peak_max_samples), np.nan)
        self.ptt = np.full((self.P, peak_max_samples), np.nan)

    @staticmethod
    def _push(ring, v):
        ring[..., :-1] = ring[..., 1:]
        ring[..., -1] = v

    def push_sample(self, ts, samples):
        """Steps 5-10 for pre-sampled ROI values (signals-only configs)."""
        self._push(self.t, ts)
        self._push(self.raw, np.asarray(samples, dtype=float))
        out = dict(samples=np.asarray(samples, dtype=float))
        proc = [preprocess(self.t, self.raw[r], self.methods, **self.kw) for r in range(self.R)]
        spec = [spectrum(px, py, self.transform, **self.kw) for px, py in proc]
        pk = [peak(f, m) for f, m in spec]
        corr = [xcorr(proc[a][0], proc[a][1], proc[b][1]) for a, b in itertools.combinations(range(self.R), 2)]
        ck = [peak(l, c) for l, c in corr]
        self._push(self.bpm_t, ts)
        self._push(self.bpm, np.array([f * 60 for f, _, _ in pk]))
        self._push(self.ptt, np.array([t * 1000 for t, _, _ in ck]))
        out.update(proc_x=[p[0] for p in proc], proc_y=[p[1] for p in proc],
                   freqs=[s[0] for s in spec], mags=[s[1] for s in spec],
                   lags=[c[0] for c in corr], corr=[c[1] for c in corr],
                   bpm=self.bpm[:, -1].copy(), ptt=self.ptt[:, -1].copy(),
                   peak_idx=np.array([i for _, _, i in pk]), peak_mag=np.array([m for _, m, _ in pk]),
                   lag_idx=np.array([i for _, _, i in ck]), lag_corr=np.array([m for _, m, _ in ck]))
        return out

    def process(self, frame, ts, rois):
        """Steps 2-10: `rois` = list of R Location 6-tuples (output of calc_roi)."""
        for r in range(self.R):
            self.roi_hist[r, :-1] = self.roi_hist[r, 1:]
            self.roi_hist[r, -1] = np.asarray(rois[r], dtype=float)
        boxes = [smooth_roi(self.roi_hist[r]) for r in range(self.R)]
        samples = [roi_sample(frame, b, self.channel) for b in boxes]
        out = self.push_sample(ts, samples)
        out['boxes'] = boxes
        return out
