"""Drop-in twin of the reference's `signal_processor` module (signal_processor.py:23-318) on top of the
B200 kernels: same enums, constants, `SignalStore`, and `SignalProcessor` constructor / methods, so the
reference's drivers (`bp.py`, `pbp.py`) and `drawer.py` run unmodified with this directory first on sys.path.

Every number is produced by libbpv (CUDA, sm_100a) through `bpv.ops` / `bpv.engine`; this file only moves
arguments and results.  There is no CPU fallback: without the built library or a CUDA device the first call
raises.  CUDA is initialised lazily inside the first call, never at import (pbp.py constructs the processor in
a child process, pbp.py:38).
"""
from __future__ import annotations

import enum
import itertools
import math
import typing

import numpy as np

import model
import profiler
import roi
import signal_data

if typing.TYPE_CHECKING:
    import inference_runner
    import video_reader


class SignalColorChannel(enum.Enum):
    GREEN = enum.auto()
    CHROM_GREEN = enum.auto()


class SignalProcessingMethod(enum.Enum):
    DIFF_1 = enum.auto()
    DIFF_2 = enum.auto()
    INTERP_LINEAR = enum.auto()
    INTERP_CUBIC = enum.auto()
    DETREND_CONST = enum.auto()
    DETREND_LINEAR = enum.auto()
    FILTER_BUTTER = enum.auto()
    FILTER_FIR = enum.auto()


class SignalSpectrumTransform(enum.Enum):
    DFT_RFFT = enum.auto()
    PGRAM_WELCH = enum.auto()
    PGRAM_LS = enum.auto()


# defaults, signal_processor.py:45-72
SIGNAL_COLOR_CHANNEL = SignalColorChannel.GREEN
ROI_MAX_SAMPLES = 1
SIGNAL_MAX_SAMPLES = 250
PEAK_MAX_SAMPLES = 50
SIGNAL_PROCESSING_METHODS = [SignalProcessingMethod.FILTER_BUTTER]
FILTER_BUTTER_ORDER = 16
FILTER_BUTTER_MIN_BW = 0.1
FILTER_FIR_TAPS = 127
FILTER_FIR_DF = 0.3
SIGNAL_SPECTRUM_TRANSFORM = SignalSpectrumTransform.PGRAM_LS
FILTER_MIN_FREQ = 0.8
FILTER_MAX_FREQ = 4.0
SPECTRUM_MIN_MAG = 0.0
SPECTRUM_MAX_MAG = 1.0
SIGNALS_MIN_LAG = -0.5
SIGNALS_MAX_LAG = 0.5
SIGNALS_MIN_CORR = -1.0
SIGNALS_MAX_CORR = 1.0


def _code(member, base=0):
    """Enum member -> libbpv integer code (include/bpv.h); anything else passes through so the C-ABI can
    reject it with NotImplementedError exactly where the reference raises (signal_processor.py:172,185,238,268)."""
    return member.value - base if isinstance(member, enum.Enum) else int(member)


class SignalStore:
    """The seven signal groups a frame leaves behind (signal_processor.py:75-84)."""

    def __init__(self, num_signals: int, roi_max_samples: int, signal_max_samples: int, peak_max_samples: int) -> None:
        pairs = math.comb(num_signals, 2)
        group = signal_data.SignalGroup
        self.sg_roi = group(num_signals, yi=(np.nan,) * 6, s_maxlen=roi_max_samples)
        self.sg_raw = group(num_signals, s_maxlen=signal_max_samples)
        self.sg_proc = group(num_signals)
        self.sg_spec = group(num_signals)
        self.sg_corr = group(pairs)
        self.sg_bpm = group(num_signals, s_maxlen=peak_max_samples)
        self.sg_ptt = group(pairs, s_maxlen=peak_max_samples)

    def snapshot(self) -> 'SignalStore':
        """What process() returns: independent of the processor's running state (the reference deep-copies
        the whole store, signal_processor.py:313).  Only the four stateful groups need copying; the per-frame
        groups (proc / spec / corr) are rebuilt from fresh device results every frame."""
        snap = object.__new__(SignalStore)
        for name, grp in vars(self).items():
            if name in ('sg_proc', 'sg_spec', 'sg_corr'):
                setattr(snap, name, grp)
                continue
            sigs = []
            for s in grp.signals:
                c = object.__new__(signal_data.Signal)
                c.__dict__.update(s.__dict__)
                c.x, c.y, c.v, c.w = s.x.copy(), s.y.copy(), s.v.copy(), s.w.copy()
                sigs.append(c)
            g = object.__new__(signal_data.SignalGroup)
            g.signals, g.num_signals, g.range_x, g.range_y = sigs, grp.num_signals, grp.range_x, grp.range_y
            setattr(snap, name, g)
        return snap


class SignalProcessor:

    def __init__(self,
                 selected_roi_configs: list[roi.ROIConfig] | None = None,
                 roi_max_samples: int = ROI_MAX_SAMPLES,
                 signal_max_samples: int = SIGNAL_MAX_SAMPLES,
                 peak_max_samples: int = PEAK_MAX_SAMPLES,
                 *,
                 color_channel: SignalColorChannel = SIGNAL_COLOR_CHANNEL,
                 processing_methods: list[SignalProcessingMethod] | None = None,
                 spectrum_transform: SignalSpectrumTransform = SIGNAL_SPECTRUM_TRANSFORM,
                 butter_order: int = FILTER_BUTTER_ORDER,
                 butter_min_bw: float = FILTER_BUTTER_MIN_BW,
                 fir_taps: int = FILTER_FIR_TAPS,
                 fir_df: float = FILTER_FIR_DF,
                 min_freq: float = FILTER_MIN_FREQ,
                 max_freq: float = FILTER_MAX_FREQ,
                 min_mag: float = SPECTRUM_MIN_MAG,
                 max_mag: float = SPECTRUM_MAX_MAG,
                 min_lag: float = SIGNALS_MIN_LAG,
                 max_lag: float = SIGNALS_MAX_LAG,
                 min_corr: float = SIGNALS_MIN_CORR,
                 max_corr: float = SIGNALS_MAX_CORR,
                 ls_num_freqs: int | None = None,
                 device: str = 'cuda') -> None:
        self.selected_roi_configs = roi.SELECTED_ROI_CONFIGS if selected_roi_configs is None else selected_roi_configs
        self.num_signals = len(self.selected_roi_configs)
        self.roi_max_samples, self.signal_max_samples, self.peak_max_samples = roi_max_samples, signal_max_samples, peak_max_samples
        self.store = SignalStore(self.num_signals, roi_max_samples, signal_max_samples, peak_max_samples)
        self.color_channel = color_channel
        self.processing_methods = SIGNAL_PROCESSING_METHODS if processing_methods is None else processing_methods
        self.spectrum_transform = spectrum_transform
        self.butter_order, self.butter_min_bw = butter_order, butter_min_bw
        self.fir_taps, self.fir_df = fir_taps, fir_df
        self.min_freq, self.max_freq = min_freq, max_freq
        self.min_mag, self.max_mag = min_mag, max_mag
        self.min_lag, self.max_lag = min_lag, max_lag
        self.min_corr, self.max_corr = min_corr, max_corr
        self.ls_num_freqs = ls_num_freqs            # extension (BASELINE config 3); None = reference grid F = n
        self.device = device
        self._engine = None
        self._engine_key = None
        self._staging = None
        self._host_bufs = {}

    # ------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------
    def _kw(self) -> dict:
        return dict(butter_order=self.butter_order, butter_min_bw=self.butter_min_bw, fir_taps=self.fir_taps,
                    fir_df=self.fir_df, min_freq=self.min_freq, max_freq=self.max_freq, ls_num_freqs=self.ls_num_freqs or 0)

    def _methods(self) -> list[int]:
        return [_code(m) for m in self.processing_methods]

    def _get_engine(self):
        """The S=1 batched engine behind process(); created on first use (inside pbp's child process).  Its device ring
        mirrors store.sg_raw: it is rebuilt and re-seeded from the store whenever the geometry (signal_max_samples, number
        of ROIs) or the store itself has been replaced since the last frame, so edits between frames behave as they do
        in the reference, which reads everything from self.store every frame."""
        from bpv.engine import BatchedSignalProcessor
        key = (self.num_signals, int(self.signal_max_samples), id(self.store), id(self.store.sg_raw))
        if self._engine is None or self._engine_key != key:
            self.num_signals = len(self.selected_roi_configs)
            self._engine = BatchedSignalProcessor(1, self.num_signals, signal_max_samples=self.signal_max_samples,
                                                  max_frames_per_step=1, store_arrays=True, device=self.device)
            self._seed_engine(self._engine)
            self._engine_key = (self.num_signals, int(self.signal_max_samples), id(self.store), id(self.store.sg_raw))
        e = self._engine          # attributes may be edited between frames, as users of the reference do
        e.color_channel = _code(self.color_channel, 1)
        e.methods, e.transform, e.kw = self._methods(), _code(self.spectrum_transform), self._kw()
        return e

    def _seed_engine(self, eng) -> None:
        """Load the raw-sample window of the store (a restored / replaced SignalStore, or a fresh NaN-prefilled one) into
        the engine's ring: the last W samples of sg_raw become global indices 0 .. W-1."""
        import torch
        W, sigs = eng.W, list(self.store.sg_raw.signals)
        if len(sigs) != eng.R:
            raise ValueError(f'store.sg_raw holds {len(sigs)} signals, the processor has {eng.R} ROIs')
        t = np.full(W, np.nan)
        y = np.full((eng.R, W), np.nan)
        for r, sg in enumerate(sigs):
            x, v = np.asarray(sg.x, dtype=float)[-W:], np.asarray(sg.y, dtype=float)[-W:]
            if r == 0 and x.size:
                t[W - x.size:] = x
            if v.size:
                y[r, W - v.size:] = v
        eng.reset()
        eng.ring_t[0, :W] = torch.from_numpy(t).to(eng.device)
        eng.ring_y[0, :, :W] = torch.from_numpy(y).to(eng.device)
        eng.count = W

    def _stage_frame(self, frame):
        """Frame -> pinned host staging buffer the ROI kernel reads zero-copy (only ROI rows cross PCIe)."""
        import torch
        frame = np.asarray(frame)
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError('frame must be uint8 [H, W, 3] BGR')
        if self._staging is None or tuple(self._staging.shape[2:]) != frame.shape:
            self._staging = torch.empty((1, 1, *frame.shape), dtype=torch.uint8, pin_memory=True)
        np.copyto(self._staging.numpy()[0, 0], frame)
        return self._staging

    @staticmethod
    def _boxes(rois):
        from bpv import _cabi
        rows = []
        for r in rois:
            r = np.asarray(r, dtype=float)
            rows.append([_cabi.NO_BOX, 0, 0, 0] if np.isnan(r).any() else [int(v) for v in r[2:6]])
        return np.asarray(rows, dtype=np.int32)

    def _window_tensors(self, signals):
        """[Signal] (same x) -> ring-shaped device tensors (S=1, R=len, cap=W) + params factory."""
        import torch
        from bpv import ops
        xs = [np.asarray(s.x, dtype=float) for s in signals]
        W = len(xs[0])
        rt = torch.from_numpy(np.ascontiguousarray(xs[0])[None]).to(self.device)
        ry = torch.from_numpy(np.stack([np.asarray(s.y, dtype=float) for s in signals])[None]).to(self.device)
        mk = lambda R: ops.make_params(1, R, W, W, W - 1, 1, 1, self._methods(), _code(self.spectrum_transform), **self._kw())
        return rt, ry, W, mk

    @staticmethod
    def _raise_status(status):
        if (status == 2).any():
            raise ValueError('`x` must be strictly increasing sequence.')       # scipy CubicSpline, via INTERP_CUBIC
        if (status == 3).any():
            raise ValueError('filter band edges are invalid for this sampling rate (scipy.signal raises here), or the '
                             'least-squares FIR system is too ill-conditioned for the Toeplitz solver (scipy falls back to lstsq)')

    # ------------------------------------------------------------------------------------------
    # the reference's method surface
    # ------------------------------------------------------------------------------------------
    @profiler.timeit
    def calc_rois(self, model_results: 'inference_runner.InferenceResults') -> list[roi.Location]:
        """Box per selected ROI from the largest detection of its model (signal_processor.py:133-155)."""
        sources = {model.ModelType.FACE_LANDMARKER: 'face_landmarker', model.ModelType.HAND_LANDMARKER: 'hand_landmarker'}
        out = []
        for cfg in self.selected_roi_configs:
            if cfg.model_type not in sources:
                raise NotImplementedError
            detections = getattr(model_results, sources[cfg.model_type]).detections
            if len(detections) == 0:
                out.append((np.nan,) * 6)
                continue
            bbox, points = detections[0]
            anchor = np.squeeze(np.mean([points[i] for i in cfg.landmark_indices], axis=0))
            ax, ay = anchor.round().astype(int)
            bw, bh = bbox[2] - bbox[0], bbox[3] - bbox[1]
            left, top, right, bottom = cfg.relative_bbox
            out.append((ax, ay, int(round(ax + left * bw)), int(round(ay + top * bh)),
                        int(round(ax + right * bw)), int(round(ay + bottom * bh))))
        return out

    @profiler.timeit
    def make_filter(self, signal_processing_method: SignalProcessingMethod, sampling_freq: float) -> np.ndarray:
        """Band-pass design on the device (signal_processor.py:158-173): sos [order, 6] or FIR taps [fir_taps]."""
        import torch
        from bpv import ops
        p = ops.make_params(1, 1, 2, 2, 1, 1, 1, [], 1, **self._kw())
        fs = torch.tensor([float(sampling_freq)], dtype=torch.float64, device=self.device)
        if signal_processing_method is SignalProcessingMethod.FILTER_BUTTER:
            filt = ops.butter_sos_design(fs, p)[0]
        elif signal_processing_method is SignalProcessingMethod.FILTER_FIR:
            filt = ops.firls_design(fs, p)[0]
        else:
            raise NotImplementedError
        filt = filt.cpu().numpy()
        if np.isnan(filt).any():
            raise ValueError('filter band edges are invalid for this sampling rate (scipy.signal raises here)')
        return filt

    @profiler.timeit
    def sample_signal(self, frame, sroi: roi.Location) -> signal_data.YType:
        return self.sample_signals(frame, [sroi])[0]

    @profiler.timeit
    def sample_signals(self, frame, rois: list[roi.Location]) -> list[signal_data.YType]:
        """F1 on one frame (signal_processor.py:176-193)."""
        import torch
        from bpv import ops
        st = self._stage_frame(frame)
        boxes = torch.from_numpy(self._boxes(rois)[None]).to(self.device)
        val, _ = ops.roi_sample(st[0], boxes, _code(self.color_channel, 1))
        return [np.float64(v) for v in val[0].cpu().numpy()]

    @profiler.timeit
    def process_signal(self, signal_raw: signal_data.Signal) -> signal_data.Signal:
        return self._preprocess([signal_raw])[0]

    @profiler.timeit
    def process_signals(self, signals_raw) -> signal_data.SignalGroup:
        return signal_data.SignalGroup(signals=self._preprocess(list(signals_raw)))

    def _preprocess(self, sigs) -> list:
        """F2 (signal_processor.py:196-245).  Signals that share x go down in one launch."""
        from bpv import ops
        out = []
        for grp in self._group_by_x(sigs):
            rt, ry, W, mk = self._window_tensors(grp)
            if W == 0:
                out.extend((s, signal_data.Signal([], [], 0)) for s in grp)
                continue
            px, py, st = ops.window_preprocess(rt, ry, mk(len(grp)))
            self._raise_status(st.cpu().numpy())
            px, py = px.cpu().numpy()[0], py.cpu().numpy()[0]
            out.extend((s, signal_data.Signal(px[i], py[i], W)) for i, s in enumerate(grp))
        by_id = {id(s): r for s, r in out}
        return [by_id[id(s)] for s in sigs]

    @staticmethod
    def _group_by_x(sigs):
        groups = []
        for s in sigs:
            x = np.asarray(s.x, dtype=float)
            for g in groups:
                gx = np.asarray(g[0].x, dtype=float)
                if gx.shape == x.shape and np.array_equal(gx, x, equal_nan=True):
                    g.append(s)
                    break
            else:
                groups.append([s])
        return groups

    @staticmethod
    def _decided(x32, y32, idx, x_peak, y_peak):
        """float32 device arrays -> the float64 arrays of the store, made consistent with the float64 peak decision of the
        device: the winning bin carries the float64 (x, y) of the peak, and no other finite bin may beat it under numpy's
        first-maximum rule (a float32-rounded neighbour of a near-tie could).  So the returned SignalStore answers
        get_peaks() exactly as sg_bpm / sg_ptt recorded it, as the reference's store does."""
        x, y = np.asarray(x32, dtype=float).copy(), np.asarray(y32, dtype=float).copy()
        idx = int(idx)
        if 0 <= idx < y.size and np.isfinite(y_peak):
            below = np.nextafter(y_peak, -np.inf)
            head, tail = y[:idx], y[idx + 1:]
            head[head >= y_peak] = below           # earlier bins must be strictly smaller (first maximum wins)
            tail[tail > y_peak] = y_peak
            x[idx], y[idx] = x_peak, y_peak
        return x, y

    def _spectrum_signal(self, freqs, mags):
        sig = signal_data.Signal(np.asarray(freqs, dtype=float), np.asarray(mags, dtype=float), s_maxlen=len(freqs))
        sig.set_range((self.min_freq, self.max_freq), (self.min_mag, self.max_mag))
        return sig

    def _corr_signal(self, lags, corr):
        sig = signal_data.Signal(np.asarray(lags, dtype=float), np.asarray(corr, dtype=float), s_maxlen=len(lags))
        sig.set_range((self.min_lag, self.max_lag), (self.min_corr, self.max_corr))
        return sig

    @profiler.timeit
    def transform_signal(self, signal_proc: signal_data.Signal) -> signal_data.Signal:
        return self._transform([signal_proc])[0]      # keeps the (min_freq, max_freq) range it sets, as the reference

    @profiler.timeit
    def transform_signals(self, signals_proc) -> signal_data.SignalGroup:
        return signal_data.SignalGroup(signals=self._transform(list(signals_proc)))

    def _transform(self, sigs) -> list:
        """F3 (signal_processor.py:248-277); spectra come back as float32 (north_star tolerance 1e-4)."""
        from bpv import ops
        res = {}
        for grp in self._group_by_x(sigs):
            rt, ry, W, mk = self._window_tensors(grp)
            if W == 0:
                res.update({id(s): self._spectrum_signal([], []) for s in grp})
                continue
            px = rt[:, None, :].expand(1, len(grp), W).contiguous()
            o = ops.window_spectrum(px, ry, mk(len(grp)))
            nb = o['num_bins'].cpu().numpy()[0]
            f, m = o['freqs'].cpu().numpy()[0].astype(float), o['mags'].cpu().numpy()[0].astype(float)
            res.update({id(s): self._spectrum_signal(f[i, :nb[i]], m[i, :nb[i]]) for i, s in enumerate(grp)})
        return [res[id(s)] for s in sigs]

    @profiler.timeit
    def correlate_signal_pair(self, signal_a: signal_data.Signal, signal_b: signal_data.Signal) -> signal_data.Signal:
        """F4 for one pair (signal_processor.py:280-295)."""
        from bpv import ops
        rt, ry, W, mk = self._window_tensors([signal_a, signal_b])
        if W == 0:
            return self._corr_signal([], [])
        px = rt[:, None, :].expand(1, 2, W).contiguous()
        o = ops.window_xcorr(px, ry, mk(2))
        n = int(o['num_lags'].cpu().numpy()[0, 0])
        return self._corr_signal(o['lags'].cpu().numpy()[0, 0, :n].astype(float), o['corr'].cpu().numpy()[0, 0, :n].astype(float))

    @profiler.timeit
    def correlate_signals(self, signals_proc) -> signal_data.SignalGroup:
        return signal_data.SignalGroup(signals=[self.correlate_signal_pair(a, b)
                                                for a, b in itertools.combinations(list(signals_proc), 2)])

    def _to_host(self, named: dict) -> dict:
        """Device tensors -> numpy arrays with ONE synchronisation: every tensor is copied asynchronously into its own
        pinned staging buffer on the stream the step ran on (sixteen blocking .cpu() calls were a quarter of the per-frame
        time of process())."""
        import torch
        bufs = self._host_bufs
        staged = {}
        with torch.cuda.device(self.device):
            for k, t in named.items():
                b = bufs.get(k)
                if b is None or b.shape != t.shape or b.dtype != t.dtype:
                    b = bufs[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                b.copy_(t, non_blocking=True)
                staged[k] = b
            torch.cuda.current_stream(self.device).synchronize()
        return {k: b.numpy().copy() for k, b in staged.items()}     # the staging buffers are reused by the next frame

    @profiler.timeit
    def process(self, frame_data: 'video_reader.FrameData', model_results: 'inference_runner.InferenceResults') -> SignalStore:
        """One frame through the whole path (signal_processor.py:302-313) as ONE batched-engine step (S=1, T=1)."""
        import torch
        st, ts = self.store, float(frame_data.timestamp)
        st.sg_roi.add_samples(ts, self.calc_rois(model_results))
        rois = st.sg_roi.get_means(as_int=True)
        eng = self._get_engine()
        boxes = torch.from_numpy(self._boxes(rois)[None, None]).to(self.device)
        tst = torch.tensor([[ts]], dtype=torch.float64, device=self.device)
        res = eng.step(self._stage_frame(frame_data.frame), boxes, tst)
        host = self._to_host(dict(res.arrays, samples=res.samples, status=res.status, peak_freq=res.peak_freq, lag_sec=res.lag_sec,
                                  peak_idx=res.peak_idx, peak_mag=res.peak_mag, lag_idx=res.lag_idx, lag_corr=res.lag_corr))
        samples, status = host['samples'][0, 0], host['status']
        peak_f, lag_s = host['peak_freq'][0], host['lag_sec'][0]
        R, W = self.num_signals, self.signal_max_samples
        st.sg_raw.add_samples(ts, [np.float64(v) for v in samples])     # before any raise, as the reference (:307-308)
        self._raise_status(status)
        st.sg_proc = signal_data.SignalGroup(signals=[signal_data.Signal(host['proc_x'][0, r], host['proc_y'][0, r], W) for r in range(R)])
        nb = host['num_bins'][0]
        pidx, pmag = host['peak_idx'][0], host['peak_mag'][0]
        st.sg_spec = signal_data.SignalGroup(signals=[
            self._spectrum_signal(*self._decided(host['freqs'][0, r, :nb[r]], host['mags'][0, r, :nb[r]], pidx[r], peak_f[r], pmag[r]))
            for r in range(R)])
        st.sg_bpm.add_samples(ts, [f * 60 for f in peak_f])             # peak decided in float64 on the device
        pairs = math.comb(R, 2)
        nl = host['num_lags'][0] if pairs else []
        lidx, lcorr = host['lag_idx'][0], host['lag_corr'][0]
        st.sg_corr = signal_data.SignalGroup(signals=[
            self._corr_signal(*self._decided(host['lags'][0, k, :nl[k]], host['corr'][0, k, :nl[k]], lidx[k], lag_s[k], lcorr[k]))
            for k in range(pairs)])
        st.sg_ptt.add_samples(ts, [t * 1000 for t in lag_s])
        return st.snapshot()

    run = process

    def cleanup(self):
        self._engine = None
        self._staging = None
        self._host_bufs = {}
