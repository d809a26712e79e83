"""GPU: the drop-in modules (roi / signal_data / signal_processor) used exactly as bp.py uses the reference's
(signal_processor.py:302-313): golden replay through SignalProcessor.process, and the per-method surface."""
import pickle

import numpy as np
import pytest

from oracle import bpv_oracle as orc
from tests import helpers as h
from tests.test_engine_gpu import MIN_N, RESIDUE

pytestmark = pytest.mark.gpu


class _Out:
    def __init__(self, d): self.detections = d


class _Res:
    def __init__(self, f, hd): self.face_landmarker, self.hand_landmarker = _Out(f), _Out(hd)


class _Frame:
    def __init__(self, frame, ts): self.frame, self.timestamp = frame, ts


def _results(g, i):
    face, hand = [], []
    if g['present'][i, 0]:
        pts = np.zeros((478, 2), np.int64); pts[151] = g['face_pt'][i]
        face = [(tuple(int(v) for v in g['face_bbox'][i]), pts)]
    if g['present'][i, 1]:
        pts = np.zeros((21, 2), np.int64); pts[0], pts[9] = g['hand_pts'][i, 0], g['hand_pts'][i, 1]
        hand = [(tuple(int(v) for v in g['hand_bbox'][i]), pts)]
    return _Res(face, hand)


@pytest.mark.parametrize('name', ['c1_butter_ls', 'c2_detrend_fir_welch', 'c4_cubic_butter_ls', 'lin_const_fir_dft', 'diff2_welch'])
def test_process_golden_replay(name):
    import signal_processor as sp
    channel, methods, transform, window, n, fps, irregular, p_none, roi_ms, kw = h.CASES[name]
    g = h.load_case(name)
    frames = h.case_frames(g)
    proc = sp.SignalProcessor(None, roi_ms, window, 50, color_channel=sp.SignalColorChannel[channel],
                              processing_methods=[sp.SignalProcessingMethod[m] for m in methods],
                              spectrum_transform=sp.SignalSpectrumTransform[transform], **kw)
    ls = transform == 'PGRAM_LS'
    full_at = set(int(i) for i in g['full_at'])
    prev = None
    for i in range(n):
        store = proc.run(_Frame(frames[i], float(g['ts'][i])), _results(g, i))
        boxes = np.array([np.asarray(b, dtype=float) for b in store.sg_roi.get_means(as_int=True)])
        assert h.same(boxes, g['boxes'][i]), (name, i)                              # calc_rois + ROI smoothing: exact
        assert h.same([s.y[-1] for s in store.sg_raw], g['raw'][i]), (name, i)      # ROI samples: bit-exact
        nvalid = np.isfinite(g['raw'][:i + 1][-window:]).sum(axis=0)
        joint = int(np.isfinite(g['raw'][:i + 1][-window:]).all(axis=1).sum())
        rawmax = np.nanmax(np.abs(g['raw'][:i + 1][-window:]), axis=0, initial=1.0)
        resid = [not np.isfinite(s.y).any() or np.nanmax(np.abs(s.y)) < RESIDUE * rawmax[r] for r, s in enumerate(store.sg_proc)]
        bpm = [s.y[-1] for s in store.sg_bpm]
        for r in range(2):
            if nvalid[r] >= MIN_N and not resid[r]:
                assert h.close(bpm[r], g['bpm'][i, r], rtol=0 if ls else 1e-12, atol_frac=0), (name, i, r)
        if joint >= MIN_N and not any(resid):
            assert h.close([s.y[-1] for s in store.sg_ptt], g['ptt'][i], rtol=1e-12, atol_frac=0), (name, i)
        assert len(store.sg_spec.signals) == 2 and len(store.sg_corr.signals) == 1
        if i in full_at:
            for r in range(2):
                assert h.close(store.sg_proc.signals[r].x, g[f'f{i}_proc_x{r}'], rtol=1e-12, atol_frac=0)
                assert h.close(store.sg_proc.signals[r].y, g[f'f{i}_proc_y{r}'], rtol=1e-4, atol_frac=1e-7, atol=2e-9)
                assert len(store.sg_spec.signals[r].x) == len(g[f'f{i}_spec_x{r}'])
            assert len(store.sg_corr.signals[0].x) == len(g[f'f{i}_corr_x0'])
        if prev is not None:
            assert prev.sg_raw.signals[0].x is not store.sg_raw.signals[0].x       # snapshots are independent
        prev = store
    pickle.loads(pickle.dumps(store))                                              # pbp.py ships it through a Manager queue


def test_method_surface_matches_oracle():
    import signal_data as sd
    import signal_processor as sp
    from bpv import synth
    rng = np.random.default_rng(8)
    proc = sp.SignalProcessor(signal_max_samples=120, color_channel=sp.SignalColorChannel.CHROM_GREEN,
                              processing_methods=[sp.SignalProcessingMethod.DETREND_LINEAR, sp.SignalProcessingMethod.FILTER_BUTTER],
                              spectrum_transform=sp.SignalSpectrumTransform.PGRAM_WELCH)
    # make_filter
    np.testing.assert_allclose(proc.make_filter(sp.SignalProcessingMethod.FILTER_BUTTER, 29.7),
                               orc.make_filter(orc.FILTER_BUTTER, 29.7), rtol=1e-9, atol=1e-300)
    ref = orc.make_filter(orc.FILTER_FIR, 29.7)
    np.testing.assert_allclose(proc.make_filter(sp.SignalProcessingMethod.FILTER_FIR, 29.7), ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    with pytest.raises(NotImplementedError):
        proc.make_filter(sp.SignalProcessingMethod.DIFF_1, 30.0)
    with pytest.raises(ValueError):
        proc.make_filter(sp.SignalProcessingMethod.FILTER_FIR, 7.5)
    # sample_signal(s) incl. a cropped (strided) frame view as video_reader.py:101 produces
    big = rng.integers(0, 256, (90, 140, 3), dtype=np.uint8)
    for frame in (big, big[:, 20:110]):
        H, W = frame.shape[:2]
        rois = [(0, 0, 5, 7, 60, 50), (0, 0, -30, -20, W + 5, H), (np.nan,) * 6, (0, 0, 10, 10, 10, 30)]
        got = proc.sample_signals(frame, rois)
        exp = [orc.roi_sample(frame, r, orc.CHROM_GREEN) for r in rois]
        assert h.same(got, exp)
        assert h.same(proc.sample_signal(frame, rois[0]), exp[0])
    # process_signal / transform_signal / correlate_signal_pair on Signals
    W = 120
    ts = synth.timestamps(rng, W, 30.0, irregular=True, drop=0.05, origin=4.0)
    ys = synth.raw_signals(rng, ts, R=2, p_nan=0.04)
    raw = [sd.Signal(list(ts), list(ys[r]), W) for r in range(2)]
    methods = [orc.DETREND_LINEAR, orc.FILTER_BUTTER]
    procd = proc.process_signals(sd.SignalGroup(signals=raw))
    for r in range(2):
        ex, ey = orc.preprocess(ts, ys[r], methods)
        assert h.close(procd.signals[r].x, ex, rtol=1e-12, atol_frac=0) and h.close(procd.signals[r].y, ey, rtol=1e-7, atol_frac=1e-7)
        one = proc.process_signal(raw[r])
        assert h.close(one.y, ey, rtol=1e-7, atol_frac=1e-7)
        spec = proc.transform_signal(procd.signals[r])
        ef, em = orc.spectrum(ex, ey, orc.PGRAM_WELCH)
        assert h.close(spec.x, ef, rtol=1e-6, atol_frac=0) and h.close(spec.y, em, rtol=1e-4, atol_frac=1e-5)
        assert spec.range_x == (proc.min_freq, proc.max_freq)            # set_range as the reference (pre-clobbering)
    ex0, ey0 = orc.preprocess(ts, ys[0], methods)
    _, ey1 = orc.preprocess(ts, ys[1], methods)
    el, ec = orc.xcorr(ex0, ey0, ey1)
    corr = proc.correlate_signals(procd)
    assert corr.num_signals == 1
    assert h.close(corr.signals[0].x, el, rtol=1e-6, atol_frac=1e-7) and h.close(corr.signals[0].y, ec, rtol=1e-4, atol_frac=1e-5)
    # errors the reference raises
    dup = ts.copy(); dup[50] = dup[49]
    proc.processing_methods = [sp.SignalProcessingMethod.INTERP_CUBIC]
    with pytest.raises(ValueError):
        proc.process_signal(sd.Signal(list(dup), list(ys[0]), W))
    proc.color_channel = 'nope'
    with pytest.raises((NotImplementedError, ValueError, TypeError)):
        proc.sample_signal(big, (0, 0, 1, 1, 5, 5))


@pytest.mark.parametrize('name', ['c1_butter_ls', 'c2_detrend_fir_welch', 'lin_const_fir_dft'])
def test_returned_store_is_self_consistent(name):
    """recomputing the peaks from the returned store (as the reference's own process() does,
    signal_processor.py:310, 312) gives exactly what sg_bpm / sg_ptt recorded — the float32 device spectra are made
    consistent with the float64 peak decision."""
    import signal_processor as sp
    channel, methods, transform, window, n, fps, irregular, p_none, roi_ms, kw = h.CASES[name]
    g = h.load_case(name)
    frames = h.case_frames(g)
    proc = sp.SignalProcessor(None, roi_ms, window, 50, color_channel=sp.SignalColorChannel[channel],
                              processing_methods=[sp.SignalProcessingMethod[m] for m in methods],
                              spectrum_transform=sp.SignalSpectrumTransform[transform], **kw)
    for i in range(n):
        store = proc.process(_Frame(frames[i], float(g['ts'][i])), _results(g, i))
        assert h.same([f * 60 for f, _ in store.sg_spec.get_peaks()], [s.y[-1] for s in store.sg_bpm]), (name, i)
        assert h.same([t * 1000 for t, _ in store.sg_corr.get_peaks()], [s.y[-1] for s in store.sg_ptt]), (name, i)


def test_store_replacement_reseeds_the_device_ring():
    """The reference reads everything from self.store each frame, so a caller may restore a pickled store (pbp.py ships
    stores between processes) or swap it: the drop-in re-seeds its device ring from the new store."""
    import signal_processor as sp
    name = 'c2_detrend_fir_welch'
    channel, methods, transform, window, n, fps, irregular, p_none, roi_ms, kw = h.CASES[name]
    g = h.load_case(name)
    frames = h.case_frames(g)
    mk = lambda: sp.SignalProcessor(None, roi_ms, window, 50, color_channel=sp.SignalColorChannel[channel],
                                    processing_methods=[sp.SignalProcessingMethod[m] for m in methods],
                                    spectrum_transform=sp.SignalSpectrumTransform[transform], **kw)
    a, b = mk(), mk()
    cut = n // 2
    full = [a.process(_Frame(frames[i], float(g['ts'][i])), _results(g, i)) for i in range(n)]
    for i in range(cut):
        st = b.process(_Frame(frames[i], float(g['ts'][i])), _results(g, i))
    c = mk()
    c.store = pickle.loads(pickle.dumps(st))
    for i in range(cut, n):
        got = c.process(_Frame(frames[i], float(g['ts'][i])), _results(g, i))
        assert h.same([s.y[-1] for s in got.sg_bpm], [s.y[-1] for s in full[i].sg_bpm]), i
        assert h.same([s.y[-1] for s in got.sg_ptt], [s.y[-1] for s in full[i].sg_ptt]), i
        assert h.same(np.asarray(got.sg_raw.signals[0].y), np.asarray(full[i].sg_raw.signals[0].y)), i


def test_exception_keeps_store_and_ring_in_step():
    """INTERP_CUBIC on duplicate timestamps raises ValueError out of process() (scipy CubicSpline in the reference); the
    reference has appended the raw sample by then (signal_processor.py:307-308), and so has the drop-in: the frames after
    the failure see the same window in the store and on the device."""
    import signal_processor as sp
    rng = np.random.default_rng(0)
    H, W = h.IMG_H, h.IMG_W
    proc = sp.SignalProcessor(None, 1, 24, 50, color_channel=sp.SignalColorChannel.GREEN,
                              processing_methods=[sp.SignalProcessingMethod.INTERP_CUBIC], spectrum_transform=sp.SignalSpectrumTransform.DFT_RFFT)
    face = [((10, 5, 50, 45), np.full((478, 2), 30, np.int64))]
    hand = [((40, 30, 70, 55), np.full((21, 2), 50, np.int64))]
    ts = [0.1, 0.2, 0.3, 0.3, 0.4, 0.5]
    raised = 0
    for i, t in enumerate(ts):
        fr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        try:
            store = proc.process(_Frame(fr, t), _Res(face, hand))
        except ValueError:
            raised += 1
        assert np.isfinite(np.asarray(proc.store.sg_raw.signals[0].y)).sum() == i + 1
        ring = proc._engine.ring_y[0, 0].cpu().numpy()
        assert np.isfinite(ring).sum() == i + 1
    assert raised >= 1
