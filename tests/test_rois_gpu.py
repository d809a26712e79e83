"""GPU: batched calc_rois + ROI smoothing against the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import bpv_oracle as orc
from tests import helpers as h

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('H', [1, 2, 3, 5])
def test_calc_rois_and_smoothing_exact(H):
    from bpv import ops, synth
    rng = np.random.default_rng(H)
    S, T, R, K = 9, 23, 3, 2
    rel = [h.REL[0], h.REL[1], (-0.05, -0.05, 0.15, 0.05)]
    npts = [1, 2, 2]
    present = (rng.uniform(size=(S, T, R)) > 0.15).astype(np.uint8)
    x0 = rng.integers(0, 1500, (S, T, R)); y0 = rng.integers(0, 900, (S, T, R))
    bbox = np.stack([x0, y0, x0 + rng.integers(1, 700, (S, T, R)), y0 + rng.integers(1, 700, (S, T, R))], axis=-1).astype(np.int32)
    points = rng.integers(0, 1920, (S, T, R, K, 2)).astype(np.int32)     # odd sums exercise the .5 rounding
    hist = torch.full((S, R, H, 6), float('nan'), dtype=torch.float64, device='cuda')
    g0 = 0
    ref_hist = np.full((S, R, H, 6), np.nan)
    for a in range(0, T, 7):                      # several calls: the history carries over
        b = min(T, a + 7)
        boxes, loc, smo = ops.calc_rois(torch.from_numpy(present[:, a:b].copy()).cuda(), torch.from_numpy(bbox[:, a:b].copy()).cuda(),
                                        torch.from_numpy(points[:, a:b].copy()).cuda(), torch.tensor(npts, dtype=torch.int32, device='cuda'),
                                        torch.tensor(rel, dtype=torch.float64, device='cuda'), hist, g0, want_locations=True)
        boxes, loc, smo = boxes.cpu().numpy(), loc.cpu().numpy(), smo.cpu().numpy()
        for s in range(S):
            for t in range(a, b):
                for r in range(R):
                    det = [(tuple(int(v) for v in bbox[s, t, r]), points[s, t, r])] if present[s, t, r] else []
                    e = orc.calc_roi(det, list(range(npts[r])), rel[r])
                    assert h.same(loc[s, t - a, r], np.asarray(e, dtype=float)), (s, t, r)
                    ref_hist[s, r, :-1] = ref_hist[s, r, 1:]
                    ref_hist[s, r, -1] = np.asarray(e, dtype=float)
                    m = np.asarray(orc.smooth_roi(ref_hist[s, r]), dtype=float)
                    assert h.same(smo[s, t - a, r], m), (s, t, r, smo[s, t - a, r], m)
                    if np.isnan(m).any():
                        assert boxes[s, t - a, r, 0] == synth.NO_BOX
                    else:
                        assert boxes[s, t - a, r].tolist() == [int(v) for v in m[2:6]]
        g0 += b - a


def test_step_detections_equals_step_with_host_boxes():
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    rng = np.random.default_rng(1)
    S, T, R, W, Hh, Ww = 4, 6, 2, 24, 60, 80
    kw = dict(signal_max_samples=W, max_frames_per_step=T, processing_methods=[orc.DETREND_LINEAR], spectrum_transform=orc.DFT_RFFT)
    e1, e2 = BatchedSignalProcessor(S, R, **kw), BatchedSignalProcessor(S, R, **kw)
    e1.set_roi_configs(h.REL, [1, 2], roi_max_samples=2)
    hist = np.full((S, R, 2, 6), np.nan)
    for step in range(5):
        ts = np.stack([(np.arange(T) + 1 + step * T) / 30.0 for _ in range(S)])
        frames = rng.integers(0, 256, (S, T, Hh, Ww, 3), dtype=np.uint8)
        present = (rng.uniform(size=(S, T, R)) > 0.1).astype(np.uint8)
        bbox = np.tile(np.array([20, 10, 60, 50], np.int32), (S, T, R, 1)) + rng.integers(-3, 4, (S, T, R, 4)).astype(np.int32)
        points = (np.array([40, 30], np.int32) + rng.integers(-4, 5, (S, T, R, 2, 2))).astype(np.int32)
        r1, boxes = e1.step_detections(torch.from_numpy(frames).cuda(), torch.from_numpy(present).cuda(), torch.from_numpy(bbox).cuda(),
                                       torch.from_numpy(points).cuda(), torch.from_numpy(ts).cuda())
        hb = np.empty((S, T, R, 4), np.int32)
        for s in range(S):
            for t in range(T):
                for r in range(R):
                    det = [(tuple(int(v) for v in bbox[s, t, r]), points[s, t, r])] if present[s, t, r] else []
                    hist[s, r, :-1] = hist[s, r, 1:]
                    hist[s, r, -1] = np.asarray(orc.calc_roi(det, [0] if r == 0 else [0, 1], h.REL[r]), dtype=float)
                    m = np.asarray(orc.smooth_roi(hist[s, r]), dtype=float)
                    hb[s, t, r] = [synth.NO_BOX, 0, 0, 0] if np.isnan(m).any() else [int(v) for v in m[2:6]]
        assert np.array_equal(boxes.cpu().numpy(), hb)
        r2 = e2.step(torch.from_numpy(frames).cuda(), torch.from_numpy(hb).cuda(), torch.from_numpy(ts).cuda())
        assert torch.equal(torch.nan_to_num(r1.samples, nan=-1), torch.nan_to_num(r2.samples, nan=-1))
        assert torch.equal(r1.peak_idx, r2.peak_idx) and torch.equal(r1.lag_idx, r2.lag_idx)


@pytest.mark.parametrize('H', [1, 5, 7, 8, 13, 50])
def test_running_means_match_numpy_nanmean(H):
    """sg_bpm / sg_ptt histories + get_means() on the device: bit-exact with np.nanmean over the
    deque in chronological order, and its half-to-even integer round."""
    from bpv import ops
    rng = np.random.default_rng(H)
    S, C = 7, 3
    ring = torch.full((S, C, H), float('nan'), dtype=torch.float64, device='cuda')
    hist = np.full((S, C, H), np.nan)
    g0 = 0
    for T in (1, 4, 9, 30, 2):
        vals = rng.uniform(0.7, 4.0, (S, T, C))
        vals[rng.uniform(size=vals.shape) < 0.2] = np.nan
        vals[0, :, 0] = np.round(vals[0, :, 0] * 2) / 2 / 60          # exact .5 means now and then
        mean, mean_int = ops.running_mean(ring, g0, torch.from_numpy(vals).cuda(), 60.0)
        mean, mean_int = mean.cpu().numpy(), mean_int.cpu().numpy()
        for t in range(T):
            hist[:, :, :-1] = hist[:, :, 1:]
            hist[:, :, -1] = vals[:, t] * 60
            for s in range(S):
                for c in range(C):
                    y = hist[s, c]
                    if np.isfinite(y).any():
                        m = np.squeeze(np.nanmean(y, axis=0))
                        assert mean[s, t, c] == m, (H, T, t, s, c, mean[s, t, c], m)
                        assert mean_int[s, t, c] == m.round()
                    else:
                        assert np.isnan(mean[s, t, c]) and np.isnan(mean_int[s, t, c])
        g0 += T


def test_engine_tracks_running_means():
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    rng = np.random.default_rng(3)
    S, W, T, R, Hm = 3, 40, 5, 2, 6
    eng = BatchedSignalProcessor(S, R, signal_max_samples=W, max_frames_per_step=T, processing_methods=[orc.DETREND_CONST],
                                 spectrum_transform=orc.PGRAM_LS, peak_max_samples=Hm)
    n = 60
    ts = np.stack([synth.timestamps(rng, n, 30.0) for _ in range(S)])
    ys = np.stack([synth.raw_signals(rng, ts[s]).T for s in range(S)])
    bpm_hist = np.full((S, R, Hm), np.nan)
    for g0 in range(0, n, T):
        res = eng.step_signals(torch.from_numpy(ys[:, g0:g0 + T].copy()).cuda(), torch.from_numpy(ts[:, g0:g0 + T].copy()).cuda())
        bpm = res.bpm.cpu().numpy().reshape(S, T, R)
        mb = res.means['mean_bpm_int'].cpu().numpy().reshape(S, T, R)
        for t in range(T):
            bpm_hist[:, :, :-1] = bpm_hist[:, :, 1:]
            bpm_hist[:, :, -1] = bpm[:, t]
            for s in range(S):
                for r in range(R):
                    y = bpm_hist[s, r]
                    exp = np.nanmean(y).round() if np.isfinite(y).any() else np.nan
                    assert (np.isnan(exp) and np.isnan(mb[s, t, r])) or mb[s, t, r] == exp
        assert res.means['mean_ptt_int'].shape == (S * T, 1)
