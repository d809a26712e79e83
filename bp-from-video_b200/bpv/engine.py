"""BatchedSignalProcessor — the reference's SignalProcessor.process() (signal_processor.py:302-313)
vectorised over S independent video streams on one GPU.

Per step (T frames per stream):
  1. F1  ROI sampling of S*T frames                       (sample_signals, :306)
  2.     push samples + timestamps into the device rings   (sg_raw.add_samples, :307)
  3. F2  preprocessing of every window job                 (process_signals, :308)
  4. F3  spectrum + heart-rate peak                        (transform_signals + get_peaks, :309-310)
  5. F4  pairwise xcorr + PTT lag peak                     (correlate_signals + get_peaks, :311-312)
A window job is one evaluation of a stream's sliding window.  `windows='every_frame'` evaluates the
window after every pushed frame, exactly as the reference does; `windows='last'` only after the last
frame of the step.  All state lives in HBM: ring_t f64 [S, cap], ring_y f64 [S, R, cap].

ROI boxes are inputs (calc_rois / detection are upstream of the path).  torch is used for
device allocations and the stream handle only; every computation is a libbpv kernel.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import torch

from . import _cabi, ops

EVERY_FRAME, LAST = 'every_frame', 'last'


@dataclass
class StepResult:
    """Device tensors for the J = S * jobs_per_stream window jobs of one step (job = s*jobs_per_stream + j)."""
    jobs_per_stream: int
    samples: torch.Tensor | None          # f64 [S, T, R]   raw ROI samples of this step
    peak_freq: torch.Tensor               # f64 [J, R]      Hz (NaN = none)
    peak_idx: torch.Tensor                # i32 [J, R]
    peak_mag: torch.Tensor                # f64 [J, R]
    lag_sec: torch.Tensor                 # f64 [J, P]
    lag_idx: torch.Tensor                 # i32 [J, P]
    lag_corr: torch.Tensor                # f64 [J, P]
    status: torch.Tensor                  # i32 [J, R]
    arrays: dict = field(default_factory=dict)   # proc_x/proc_y/freqs/mags/lags/corr/num_bins/num_lags when stored
    means: dict = field(default_factory=dict)    # mean_bpm, mean_bpm_int [J, R], mean_ptt, mean_ptt_int [J, P] when tracked

    @property
    def bpm(self):                        # signal_processor.py:310  f * 60
        return self.peak_freq * 60

    @property
    def ptt_ms(self):                     # signal_processor.py:312  t * 1000
        return self.lag_sec * 1000

    def packed(self) -> torch.Tensor:
        """[J, 2R + 2P] float64 record (bpm, ptt_ms, peak_idx, lag_idx) — what the multi-GPU gather moves."""
        return ops.pack_records(self.peak_freq, self.lag_sec, self.peak_idx, self.lag_idx)


class BatchedSignalProcessor:
    def __init__(self, num_streams: int, num_rois: int = 2, *, signal_max_samples: int = 250,
                 max_frames_per_step: int = 1, color_channel: int = _cabi.GREEN, processing_methods=(_cabi.FILTER_BUTTER,),
                 spectrum_transform: int = _cabi.PGRAM_LS, butter_order: int = 16, butter_min_bw: float = 0.1,
                 fir_taps: int = 127, fir_df: float = 0.3, min_freq: float = 0.8, max_freq: float = 4.0,
                 ls_num_freqs: int | None = None, windows: str = EVERY_FRAME, store_arrays: bool = False,
                 device: str | torch.device = 'cuda', roi_pixels_hint: int = 0, peak_max_samples: int = 0):
        _cabi.lib()  # fail loudly if the CUDA library is missing: there is no CPU fallback
        if not torch.cuda.is_available():
            raise _cabi.BpvError('BatchedSignalProcessor needs a CUDA device (no CPU fallback)')
        assert windows in (EVERY_FRAME, LAST)
        self.S, self.R, self.P = int(num_streams), int(num_rois), math.comb(int(num_rois), 2)
        self.W = int(signal_max_samples)
        self.Tmax = int(max_frames_per_step)
        self.cap = self.W + self.Tmax
        self.color_channel = int(color_channel)
        self.methods = [int(m) for m in processing_methods]
        self.transform = int(spectrum_transform)
        self.kw = dict(butter_order=butter_order, butter_min_bw=butter_min_bw, fir_taps=fir_taps, fir_df=fir_df,
                       min_freq=min_freq, max_freq=max_freq, ls_num_freqs=int(ls_num_freqs or 0))
        self.windows, self.store_arrays = windows, bool(store_arrays)
        self.roi_pixels_hint = int(roi_pixels_hint)
        self.device = torch.device(device)
        dev, f64 = self.device, torch.float64
        self.ring_t = torch.full((self.S, self.cap), float('nan'), dtype=f64, device=dev)
        self.ring_y = torch.full((self.S, self.R, self.cap), float('nan'), dtype=f64, device=dev)
        self.count = 0  # samples pushed per stream so far (global index of the next sample)
        Jmax = self.S * (self.Tmax if windows == EVERY_FRAME else 1)
        self._Jmax = Jmax
        self._samples = torch.empty((self.S, self.Tmax, self.R), dtype=f64, device=dev)
        self._proc_x = torch.empty((Jmax, self.R, self.W), dtype=f64, device=dev)
        self._proc_y = torch.empty((Jmax, self.R, self.W), dtype=f64, device=dev)
        self._status = torch.empty((Jmax, self.R), dtype=torch.int32, device=dev)
        p = self._params(0, 1, self.Tmax if windows == EVERY_FRAME else 1)
        self._mb = ops.max_bins(p)
        need = max(_cabi.lib().bpv_window_workspace_bytes(p), _cabi.lib().bpv_spectrum_workspace_bytes(p, self._mb), 16)
        self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._spec = self._xc = None
        self.launches_per_step = self._extra_launches = 0
        # sg_bpm / sg_ptt histories and their running means on the device (0 = off)
        self.peak_max_samples = int(peak_max_samples)
        if self.peak_max_samples:
            self._bpm_ring = torch.full((self.S, self.R, self.peak_max_samples), float('nan'), dtype=f64, device=dev)
            self._ptt_ring = torch.full((self.S, max(self.P, 1), self.peak_max_samples), float('nan'), dtype=f64, device=dev)
            self._mean_count = 0

    # ------------------------------------------------------------------------------------------
    def _params(self, head0: int, head_step: int, jobs: int) -> _cabi.WindowParams:
        return ops.make_params(self.S, self.R, self.cap, self.W, head0, head_step, jobs, self.methods, self.transform, **self.kw)

    def step(self, frames: torch.Tensor, boxes: torch.Tensor, timestamps: torch.Tensor, view=None, nv12_size=None,
             resize_to=None) -> StepResult:
        """frames uint8 [S, T, H, W, 3] (HBM, or pinned host memory: the ROI kernel then reads the ROI rows
        straight over PCIe), boxes int32 [S, T, R, 4] (device), timestamps float64 [S, T] (device).
        view = (view_w, view_h, left, flip_horizontally): the boxes are expressed in the reference VideoReader's
        cropped / mirrored view of the decoded frames (video_reader.py:97-103) and are mapped back onto `frames`
        on the device; nothing is copied.
        nv12_size = (H, W): frames are NV12 decoder buffers uint8 [S, T, 3H/2, pitch]; the ROI is sampled from the planes
        with OpenCV's integer BT.601 conversion, i.e. as from the BGR frame cv2.VideoCapture would have produced."""
        S, T = frames.shape[:2]
        assert S == self.S and 1 <= T <= self.Tmax and boxes.shape == (S, T, self.R, 4)
        self._extra_launches = 0
        if view is not None:
            boxes = ops.view_boxes(boxes.contiguous(), *view)
            self._extra_launches = 1
        if resize_to is not None:
            # boxes live in the frame cv2.resize(frame, (dst_w, dst_h)) would produce (video_reader.py:95-96); pixels are
            # generated on the fly with OpenCV's integer bilinear arithmetic, the resized frame is never materialised
            dst_h, dst_w = resize_to
            val, _ = ops.roi_sample_resized(frames.view(S * T, *frames.shape[2:]), dst_h, dst_w,
                                            boxes.reshape(S * T, self.R, 4).contiguous(), self.color_channel)
            return self.step_signals(val.view(S, T, self.R), timestamps, _count_roi=True)
        if nv12_size is not None:
            H, W = nv12_size
            val, _ = ops.roi_sample_nv12(frames.view(S * T, *frames.shape[2:]), H, W, boxes.reshape(S * T, self.R, 4).contiguous(),
                                         self.color_channel)
            return self.step_signals(val.view(S, T, self.R), timestamps, _count_roi=True)
        samples = self._samples[:, :T]
        if T != self.Tmax:
            samples = torch.empty((S, T, self.R), dtype=torch.float64, device=self.device)
        ops.roi_sample(frames.view(S * T, *frames.shape[2:]), boxes.view(S * T, self.R, 4), self.color_channel,
                       roi_pixels_hint=self.roi_pixels_hint, out_value=samples.view(S * T, self.R))
        return self.step_signals(samples, timestamps, _count_roi=True)

    # ------------------------------------------------------------------------------------------
    # calc_rois + ROI smoothing on the device for batched landmark tensors
    def set_roi_configs(self, relative_bboxes, num_points, roi_max_samples: int = 1):
        """relative_bboxes [R][4] (left, top, right, bottom) and the number of anchor landmarks per ROI, as in
        roi.ROIConfig (roi.py:8-13); roi_max_samples = length of the smoothing history (signal_processor.py:47)."""
        assert len(relative_bboxes) == self.R and len(num_points) == self.R
        self._rel = torch.tensor(relative_bboxes, dtype=torch.float64, device=self.device).contiguous()
        self._npts = torch.tensor(num_points, dtype=torch.int32, device=self.device)
        self._roi_hist = torch.full((self.S, self.R, int(roi_max_samples), 6), float('nan'), dtype=torch.float64, device=self.device)
        self._roi_count = 0

    def step_detections(self, frames, present, bbox, points, timestamps, want_locations: bool = False):
        """One step from detector outputs instead of boxes: present u8 [S,T,R], bbox i32 [S,T,R,4] (largest detection
        of the ROI's model), points i32 [S,T,R,K,2] (the landmarks the ROI config selects).  Returns (StepResult, boxes
        [, locations, smoothed])."""
        out = ops.calc_rois(present, bbox, points, self._npts, self._rel, self._roi_hist, self._roi_count, want_locations)
        self._roi_count += present.shape[1]
        boxes = out[0] if want_locations else out
        res = self.step(frames, boxes, timestamps)
        return (res, *out) if want_locations else (res, boxes)

    def roi_samples(self, frames: torch.Tensor, boxes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """F1 only, on the current stream: float64 [S, T, R] samples of frames uint8 [S, T, H, W, 3].  Lets a caller
        overlap the ROI sampling of the next batch (e.g. zero-copy from pinned host memory, PCIe bound) with the
        window pipeline of the current one: run this on a side stream, then `step_signals` on the main stream."""
        S, T = frames.shape[:2]
        assert S == self.S and boxes.shape == (S, T, self.R, 4)
        if out is None:
            out = torch.empty((S, T, self.R), dtype=torch.float64, device=self.device)
        ops.roi_sample(frames.view(S * T, *frames.shape[2:]), boxes.view(S * T, self.R, 4), self.color_channel,
                       roi_pixels_hint=self.roi_pixels_hint, out_value=out.view(S * T, self.R))
        return out

    def step_signals(self, samples: torch.Tensor, timestamps: torch.Tensor, _count_roi: bool = False) -> StepResult:
        """Signals-only entry: samples float64 [S, T, R] (already ROI-sampled), timestamps float64 [S, T]."""
        S, T, R = samples.shape
        assert S == self.S and R == self.R and 1 <= T <= self.Tmax and timestamps.shape == (S, T)
        ops.ring_push(self.ring_t, self.ring_y, self.count, timestamps.contiguous(), samples.contiguous())
        g0 = self.count
        self.count += T
        jobs = T if self.windows == EVERY_FRAME else 1
        head0 = g0 if self.windows == EVERY_FRAME else g0 + T - 1
        p = self._params(head0, 1, jobs)
        J = S * jobs
        px, py, st = self._proc_x[:J], self._proc_y[:J], self._status[:J]
        ops.window_preprocess(self.ring_t, self.ring_y, p, px, py, st, workspace=self._ws)
        sp = ops.window_spectrum(px, py, p, store=self.store_arrays, workspace=self._ws,
                                 out=None if self._spec is None or self._spec['peak_idx'].shape[0] != J else self._spec)
        xc = ops.window_xcorr(px, py, p, store=self.store_arrays,
                              out=None if self._xc is None or self._xc['lag_idx'].shape[0] != J else self._xc)
        self._spec, self._xc = sp, xc
        n_pre = 1 + sum(1 for m in set(self.methods) if m in (_cabi.FILTER_BUTTER, _cabi.FILTER_FIR))
        n_spec = 2 if self.transform == _cabi.PGRAM_LS else 1
        if (self.transform == _cabi.PGRAM_WELCH and not self.store_arrays and 256 <= self.W <= 383
                and os.environ.get('BPV_WELCH_TC', '') == '1'):
            n_spec = 2                  # welch_tc_kernel (tensor cores) + welch_warp_kernel for the windows it flags
        if self.transform == _cabi.DFT_RFFT and 16 <= self.W <= 2048:
            env = os.environ.get('BPV_DFT_TC')
            interp = any(m in (_cabi.INTERP_LINEAR, _cabi.INTERP_CUBIC) for m in self.methods)
            if (env[:1] == '1') if env is not None else interp:
                n_spec = 3              # dft_tc_kernel + dft_peak_kernel + spectrum_dense_kernel for the flagged windows
        self.launches_per_step = (1 + self._extra_launches if _count_roi else 0) + 1 + n_pre + n_spec + (1 if self.P else 0)
        arrays = {}
        if self.store_arrays:
            arrays = dict(proc_x=px, proc_y=py, freqs=sp['freqs'], mags=sp['mags'], num_bins=sp['num_bins'],
                          lags=xc['lags'], corr=xc['corr'], num_lags=xc['num_lags'])
        means = {}
        if self.peak_max_samples:
            mb, mbi = ops.running_mean(self._bpm_ring, self._mean_count, sp['peak_freq'].view(S, jobs, self.R), 60.0)
            means = dict(mean_bpm=mb.view(J, self.R), mean_bpm_int=mbi.view(J, self.R))
            if self.P:
                mp, mpi = ops.running_mean(self._ptt_ring, self._mean_count, xc['lag_sec'].view(S, jobs, self.P), 1000.0)
                means.update(mean_ptt=mp.view(J, self.P), mean_ptt_int=mpi.view(J, self.P))
            self._mean_count += jobs
            self.launches_per_step += 2 if self.P else 1
        return StepResult(jobs, samples, sp['peak_freq'], sp['peak_idx'], sp['peak_mag'], xc['lag_sec'], xc['lag_idx'],
                          xc['lag_corr'], st, arrays, means)

    def reset(self):
        self.ring_t.fill_(float('nan'))
        self.ring_y.fill_(float('nan'))
        self.count = 0
