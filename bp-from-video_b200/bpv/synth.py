"""Synthetic workloads of the shapes BASELINE.json names.

numpy only (used by tests, the golden generator, smoke() and bench.py).  Everything is seeded
through `np.random.default_rng`, so the GPU box regenerates identical inputs from a seed.
"""
from __future__ import annotations

import numpy as np

NO_BOX = np.iinfo(np.int32).min  # x0 sentinel: "no detection" (the reference's NaN Location)

# roi.py:24-28 relative boxes of the two selected ROIs (forehead, palm)
REL_FOREHEAD = (-0.00, -0.10, 0.20, 0.05)
REL_PALM = (-0.10, -0.10, 0.10, 0.10)
SKIN_BGR = (110, 140, 190)


def timestamps(rng, n, fps, irregular=False, drop=0.0, origin=None):
    """Regular: (i+1)/fps.  Irregular: cumsum(1/fps*(1+U(-0.3,0.3))) with `drop` of the frames
    removed; per-stream origin U(0,1000) s when `origin` is None and irregular."""
    if not irregular:
        return (np.arange(n) + 1.0) / fps + (0.0 if origin is None else origin)
    m = int(np.ceil(1.25 * n / max(1e-9, 1.0 - drop))) + 32
    t = np.cumsum((1.0 / fps) * (1.0 + rng.uniform(-0.3, 0.3, m)))
    keep = rng.uniform(size=m) >= drop
    t = t[keep][:n]
    assert t.size == n
    return t + (rng.uniform(0, 1000) if origin is None else origin)


def roi_boxes(rng, n, H, W, jitter=2, p_none=0.01, p_oob=0.001):
    """Boxes [n, 2, 4] int32 (x0,y0,x1,y1) for forehead + palm as the reference's calc_rois
    yields them for a face bbox 0.25W x 0.4H centred (0.45W, 0.3H) and a hand bbox
    0.2W x 0.33H at (0.7W, 0.75H); +-jitter px per frame, p_none missing detections
    (x0 = NO_BOX), p_oob boxes pushed partly out of the frame."""
    out = np.empty((n, 2, 4), np.int32)
    specs = [((0.45 * W, 0.30 * H), (0.25 * W, 0.40 * H), REL_FOREHEAD),
             ((0.70 * W, 0.75 * H), (0.20 * W, 0.33 * H), REL_PALM)]
    for r, ((cx, cy), (bw, bh), (l, t, rr, b)) in enumerate(specs):
        ax = np.rint(cx + rng.integers(-jitter, jitter + 1, n)).astype(np.int64)
        ay = np.rint(cy + rng.integers(-jitter, jitter + 1, n)).astype(np.int64)
        out[:, r, 0] = np.rint(ax + l * bw)
        out[:, r, 1] = np.rint(ay + t * bh)
        out[:, r, 2] = np.rint(ax + rr * bw)
        out[:, r, 3] = np.rint(ay + b * bh)
        oob = rng.uniform(size=n) < p_oob
        out[oob, r, 0] += W // 2
        out[oob, r, 2] += W // 2
        none = rng.uniform(size=n) < p_none
        out[none, r, :] = 0
        out[none, r, 0] = NO_BOX
    return out


def frames(rng, ts, H, W, f_pulse=1.2, amp=2.0, noise=8, delay_s=0.03, boxes=None):
    """uint8 [n,H,W,3] BGR: skin-tone base + amp*sin(2*pi*f*t) on G (palm half of the frame
    delayed by delay_s) + uniform noise +-noise, clipped."""
    n = len(ts)
    f = rng.integers(-noise, noise + 1, (n, H, W, 3), dtype=np.int16)
    f += np.asarray(SKIN_BGR, np.int16)
    pulse_face = amp * np.sin(2 * np.pi * f_pulse * np.asarray(ts))
    pulse_hand = amp * np.sin(2 * np.pi * f_pulse * (np.asarray(ts) - delay_s))
    half = H // 2
    f[:, :half, :, 1] += np.rint(pulse_face).astype(np.int16)[:, None, None]
    f[:, half:, :, 1] += np.rint(pulse_hand).astype(np.int16)[:, None, None]
    return np.clip(f, 0, 255).astype(np.uint8)


def raw_signals(rng, ts, R=2, f_pulse=None, dc=120.0, ac=0.5, noise=0.15, delay_s=0.03, p_nan=0.0):
    """Signals-only workloads (configs 3, 5@1GPU): ROI-mean-like samples [R, n] float64 with
    DC ~ 120, pulse amplitude ~0.5, white noise, slow drift; ROI r delayed by r*delay_s."""
    ts = np.asarray(ts, dtype=np.float64)
    f_pulse = rng.uniform(0.8, 3.0) if f_pulse is None else f_pulse
    y = np.empty((R, ts.size))
    for r in range(R):
        tt = ts - ts[0] - r * delay_s
        y[r] = (dc + 3.0 * r + ac * np.sin(2 * np.pi * f_pulse * tt) + 0.2 * ac * np.sin(4 * np.pi * f_pulse * tt + 0.7)
                + 0.3 * np.sin(2 * np.pi * 0.11 * tt + r) + noise * rng.standard_normal(ts.size))
    if p_nan > 0:
        y[rng.uniform(size=y.shape) < p_nan] = np.nan
    return y


def detections(rng, n, H, W, jitter=2, p_none=0.01):
    """Detector outputs per frame, in the form the reference's calc_rois consumes (signal_processor.py:133-155;
    inference_runner.py:125-132): face bbox [n,4] + its landmark 151 [n,2], hand bbox [n,4] + its landmarks 0 and 9
    [n,2,2], present [n,2].  Geometry as roi_boxes: face bbox 0.25W x 0.4H around (0.45W, 0.3H), hand bbox 0.2W x 0.33H
    around (0.7W, 0.75H), anchors jittered by +-jitter px, p_none missing detections."""
    out = {}
    for name, (cx, cy), (bw, bh) in (('face', (0.45 * W, 0.30 * H), (0.25 * W, 0.40 * H)),
                                     ('hand', (0.70 * W, 0.75 * H), (0.20 * W, 0.33 * H))):
        ax = np.rint(cx + rng.integers(-jitter, jitter + 1, n)).astype(np.int64)
        ay = np.rint(cy + rng.integers(-jitter, jitter + 1, n)).astype(np.int64)
        out[name + '_bbox'] = np.stack([np.rint(ax - bw / 2), np.rint(ay - bh / 2), np.rint(ax - bw / 2) + np.rint(bw),
                                        np.rint(ay - bh / 2) + np.rint(bh)], axis=1).astype(np.int64)
        out[name + '_pt'] = np.stack([ax, ay], axis=1)
    out['present'] = rng.uniform(size=(n, 2)) >= p_none
    return out


class _ModelOut:
    def __init__(self, detections):
        self.detections = detections


class ModelResults:
    """Stand-in for inference_runner.InferenceResults: the two attributes calc_rois reads (signal_processor.py:136-139)."""

    def __init__(self, det, i):
        face, hand = [], []
        if det['present'][i, 0]:
            pts = np.zeros((478, 2), np.int64)
            pts[151] = det['face_pt'][i]
            face = [(tuple(int(v) for v in det['face_bbox'][i]), pts)]
        if det['present'][i, 1]:
            pts = np.zeros((21, 2), np.int64)
            pts[0] = pts[9] = det['hand_pt'][i]
            hand = [(tuple(int(v) for v in det['hand_bbox'][i]), pts)]
        self.face_landmarker, self.hand_landmarker = _ModelOut(face), _ModelOut(hand)


class FrameData:
    """Stand-in for video_reader.FrameData: the two attributes process() reads (signal_processor.py:304-307)."""

    def __init__(self, frame, timestamp):
        self.frame, self.timestamp = frame, timestamp
