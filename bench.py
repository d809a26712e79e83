#!/usr/bin/env python
"""bench.py — the judged benchmark of the B200 signal path (contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1] ("c2"): 256 streams x 1080p, CHROM_GREEN ROI sampling,
[DETREND_LINEAR, FILTER_FIR], Welch HR, W=300, T=32 frames per stream resident in HBM per step (51 GB,
far larger than the 126 MB L2, so nothing is cached between steps).  One step = the whole hot path over
one batch: ROI-sample S*T frames, push, and evaluate the sliding window after EVERY frame, as the
reference's process() does (S*T window jobs -> R bpm + P ptt each).  metric = ROI-sampled frames/s.
N>1: every rank owns S streams (weak scaling, no data-path collective) and the per-stream records are
all-gathered over NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: S, H, W, T, window, channel, methods, transform, fps, kwargs
    'c2': dict(S=256, H=1080, W=1920, T=32, window=300, channel='CHROM_GREEN', methods=['DETREND_LINEAR', 'FILTER_FIR'],
               transform='PGRAM_WELCH', fps=30.0, kw={},
               desc='BASELINE configs[1]: 256 streams 1080p 30 fps, chrom-green ROI + detrend + FIR + Welch HR'),
    'c2small': dict(S=32, H=1080, W=1920, T=8, window=300, channel='CHROM_GREEN', methods=['DETREND_LINEAR', 'FILTER_FIR'],
                    transform='PGRAM_WELCH', fps=30.0, kw={}, desc='reduced c2 for quick checks (not a bench line)'),
}
METRIC, UNIT = 'roi_sampled_frames_per_s', 'frames/s'
ROI_NCU_TRAFFIC_BYTES = 336.0e6   # dram__bytes_read.sum + dram__bytes_write.sum of one c2 F1 launch (profiles/r1w_c2_summary.md)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's per-frame process() on host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One stream: prefill the window, then time `frames` calls of the reference's per-frame step."""
    import warnings
    warnings.simplefilter('ignore')
    from oracle import bpv_oracle as orc
    from bpv import synth
    wl, seed, frames = args
    rng = np.random.default_rng(seed)
    H, W, win = wl['H'], wl['W'], wl['window']
    methods = [getattr(orc, m) for m in wl['methods']]
    st = orc.OracleStream(2, 1, win, 50, getattr(orc, wl['channel']), methods, getattr(orc, wl['transform']), **wl['kw'])
    ts = synth.timestamps(rng, win + frames, wl['fps'])
    st.t[:] = ts[:win]
    st.raw[:] = synth.raw_signals(rng, ts[:win])
    pool = rng.integers(0, 256, (4, H, W, 3), dtype=np.uint8)       # 4 distinct frames, cycled
    boxes = synth.roi_boxes(rng, frames, H, W)
    t0 = time.perf_counter()
    for i in range(frames):
        rois = [(np.nan,) * 6 if boxes[i, r, 0] == synth.NO_BOX else (0, 0, *[int(v) for v in boxes[i, r]]) for r in range(2)]
        st.process(pool[i % 4], float(ts[win + i]), rois)
    return time.perf_counter() - t0


def cpu_reference(wl, frames_per_stream, procs, reps=1):
    """frames/s of the oracle port with `procs` worker processes, one stream each (streams are independent).
    Returns [(aggregate rate, slowest worker seconds)] per repetition; rates exclude process start + prefill."""
    import multiprocessing as mp
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    ctx = mp.get_context('spawn')
    out = []
    with ctx.Pool(procs) as pool:
        for rep in range(reps):
            times = pool.map(_cpu_worker, [(wl, 1000 + 97 * rep + i, frames_per_stream) for i in range(procs)])
            out.append((sum(frames_per_stream / t for t in times), max(times)))
    return out


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is Python and
    cannot travel to the GPU box) on all host cores.  Each "step" is a bounded sample; rank 0 only."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    cores = os.cpu_count() or 1
    per = max(8, int(args.ref_frames))
    warm, steps = min(args.warmup, 1), max(1, min(args.steps, 5))
    res = cpu_reference(wl, per, cores, reps=warm + steps)[warm:]
    v = float(np.median([r for r, _ in res]))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': 1e3 * float(np.median([t for _, t in res])), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'desc': wl['desc'], 'windows': 'every_frame'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{cores} streams x {per} steady-state frames each per step (window prefilled), oracle port '
                                   'of SignalProcessor.process per frame (numpy/scipy, same calls as the reference), one '
                                   'process per core'},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_frames_device(torch, S, T, H, W, seed):
    """Synthetic frames in HBM: skin-tone base + per-(stream, frame) pulse on G + uniform noise +-8."""
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    frames = torch.empty((S, T, H, W, 3), dtype=torch.uint8, device='cuda')
    base = torch.tensor([110, 140, 190], dtype=torch.int16, device='cuda')
    f_pulse = torch.rand(S, generator=g, device='cuda') * 2.2 + 0.8
    tt = (torch.arange(T, device='cuda') + 1) / 30.0
    pulse = (2.0 * torch.sin(2 * math.pi * f_pulse[:, None] * tt[None, :])).round().to(torch.int16)   # [S, T]
    for s in range(S):
        noise = torch.randint(-8, 9, (T, H, W, 3), generator=g, device='cuda', dtype=torch.int16)
        noise += base
        noise[..., 1] += pulse[s][:, None, None]
        frames[s] = noise.clamp_(0, 255).to(torch.uint8)
        del noise
    return frames


def roi_bytes(boxes_np, H, W):
    x0, y0, x1, y1 = [boxes_np[..., i].astype(np.int64) for i in range(4)]
    valid = boxes_np[..., 0] != np.iinfo(np.int32).min

    def sl(a, b, L):
        a = np.where(a < 0, np.maximum(a + L, 0), np.minimum(a, L))
        b = np.where(b < 0, np.maximum(b + L, 0), np.minimum(b, L))
        return np.maximum(b - a, 0)
    return int((3 * sl(x0, x1, W) * sl(y0, y1, H) * valid).sum())


def run_gpu(args, wl):
    import torch
    from bpv import _cabi, dist as bdist, synth
    from bpv.engine import BatchedSignalProcessor
    import torch.distributed as tdist
    rank, local, world = bdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    S, T, H, W, win = wl['S'], wl['T'], wl['H'], wl['W'], wl['window']
    methods = [getattr(_cabi, m) for m in wl['methods']]
    eng = BatchedSignalProcessor(S, 2, signal_max_samples=win, max_frames_per_step=T, color_channel=getattr(_cabi, wl['channel']),
                                 processing_methods=methods, spectrum_transform=getattr(_cabi, wl['transform']),
                                 windows=args.windows, device=dev, roi_pixels_hint=int(0.05 * W * 0.06 * H), **wl['kw'])
    rng = np.random.default_rng(1234 + rank)
    frames = make_frames_device(torch, S, T, H, W, 77 + rank)
    boxes_np = np.stack([synth.roi_boxes(rng, T, H, W) for _ in range(S)])
    boxes = torch.from_numpy(boxes_np).to(dev)
    alg_roi_bytes = roi_bytes(boxes_np, H, W)
    fps = wl['fps']
    # prefill the rings with `win` synthetic samples so every timed window is full (steady state)
    ts0 = np.stack([synth.timestamps(rng, win, fps) for _ in range(S)])
    raw0 = np.stack([synth.raw_signals(rng, ts0[s]).T for s in range(S)])            # [S, win, R]
    eng.windows, keep = 'last', eng.windows
    for a in range(0, win, T):
        eng.step_signals(torch.from_numpy(raw0[:, a:a + T].copy()).to(dev), torch.from_numpy(ts0[:, a:a + T].copy()).to(dev))
    eng.windows = keep
    t_next = [float(ts0[0, -1])]
    # timestamps of every step are synthetic inputs too: resident in HBM before the timed region ([steps, S, T] float64)
    n_steps = max(3, args.warmup) + args.steps
    ts_all = (t_next[0] + (torch.arange(n_steps * T, device=dev, dtype=torch.float64) + 1) / fps).view(n_steps, 1, T).expand(n_steps, S, T).contiguous()
    t_next[0] += n_steps * T / fps
    step_no = [0]

    def one_step():
        ts = ts_all[step_no[0]]
        step_no[0] += 1
        res = eng.step(frames, boxes, ts)
        rec = res.packed()
        if world > 1:
            rec = bdist.gather_records(rec)
        return res, rec

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    # ---- timed region: exactly K steps, CUDA events, kernel families timed with events on the same stream
    ev = lambda: torch.cuda.Event(enable_timing=True)
    fam = {k: [] for k in ('roi', 'push', 'preprocess', 'spectrum', 'xcorr')}
    from bpv import ops
    orig = {k: getattr(ops, k) for k in ('roi_sample', 'ring_push', 'window_preprocess', 'window_spectrum', 'window_xcorr')}
    names = dict(roi_sample='roi', ring_push='push', window_preprocess='preprocess', window_spectrum='spectrum', window_xcorr='xcorr')

    def timed(fn, key):
        def wrap(*a, **k):
            e0, e1 = ev(), ev()
            e0.record()
            out = fn(*a, **k)
            e1.record()
            fam[key].append((e0, e1))
            return out
        return wrap
    for k, fn in orig.items():
        setattr(ops, k, timed(fn, names[k]))
    barrier()
    start, stop = ev(), ev()
    start.record()
    for _ in range(args.steps):
        res, rec = one_step()
    stop.record()
    barrier()
    for k, fn in orig.items():
        setattr(ops, k, fn)
    ms = start.elapsed_time(stop)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        ms = float(tt.item())
    fam_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in fam.items() if v}

    # ---- e2e: the same step through the public API with HOST inputs (pinned), results read back to host.
    # Frames stay in pinned host memory and F1 reads the ROI rows straight over PCIe (zero-copy): only the
    # bytes the path needs cross the bus; boxes + timestamps are copied H2D; bpm/ptt records copied D2H.
    Te = min(T, args.e2e_frames)
    host_frames = torch.empty((S, Te, H, W, 3), dtype=torch.uint8, pin_memory=True)
    host_frames.copy_(frames[:, :Te])
    host_boxes = torch.from_numpy(boxes_np[:, :Te].copy()).pin_memory()
    host_ts = torch.empty((S, Te), dtype=torch.float64, pin_memory=True)
    host_out = torch.empty((S * (Te if args.windows == 'every_frame' else 1), rec.shape[1]), dtype=torch.float64, pin_memory=True)
    e2e_roi_bytes = roi_bytes(boxes_np[:, :Te], H, W)

    # Two-stage software pipeline over two CUDA streams: F1 of batch k+1 (PCIe bound: ROI rows read straight from
    # pinned host memory) runs on a side stream while the window pipeline of batch k runs on the main stream and its
    # records are copied back.  Every batch's inputs start in host memory and its results end in host memory inside
    # the timed region; K batches are timed from the first F1 launch to the last result landing on the host.
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    sbuf = [torch.empty((S, Te, 2), dtype=torch.float64, device=dev) for _ in range(2)]
    bbuf = [torch.empty((S, Te, 2, 4), dtype=torch.int32, device=dev) for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def launch_f1(k):
        side.wait_event(done[k % 2])                        # buffer k%2 is free once batch k-2 has consumed it
        with torch.cuda.stream(side):
            bbuf[k % 2].copy_(host_boxes, non_blocking=True)
            eng.roi_samples(host_frames, bbuf[k % 2], out=sbuf[k % 2])
            evs[k % 2].record(side)

    def finish(k):
        host_ts.copy_(torch.from_numpy((t_next[0] + (np.arange(Te) + 1) / fps)[None, :].repeat(S, 0)))
        t_next[0] += Te / fps
        main.wait_event(evs[k % 2])
        r = eng.step_signals(sbuf[k % 2], host_ts.to(dev, non_blocking=True))
        done[k % 2].record(main)
        host_out.copy_(r.packed(), non_blocking=True)
        torch.cuda.synchronize(dev) if False else main.synchronize()

    def run_e2e(K):
        launch_f1(0)
        for k in range(K):
            if k + 1 < K:
                launch_f1(k + 1)
            finish(k)

    run_e2e(4)
    barrier()
    ke = max(4, min(args.steps, 40))
    t0 = time.perf_counter()
    run_e2e(ke)
    barrier()
    e2e_s = (time.perf_counter() - t0) / ke
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    if rank != 0:
        return
    peak, peak_src = peaks()
    jobs = S * (T if args.windows == 'every_frame' else 1)
    nsig = jobs * 2
    # algorithmic bytes per launch
    alg = {
        'roi': alg_roi_bytes + S * T * 2 * (16 + 8),
        'push': S * T * 3 * 8 * 2,
        'preprocess': nsig * win * (8 + 8 + 8 + 8),          # ring t,y in + proc x,y out (float64)
        'spectrum': nsig * win * 16 + nsig * 24,
        'xcorr': jobs * (3 * win * 8 + 24),
    }
    kernels = {k: {'ms': fam_ms[k], 'share': fam_ms[k] / sum(fam_ms.values()), 'alg_bytes': alg[k],
                   'gbs': alg[k] / fam_ms[k] / 1e6, 'frac_hbm': alg[k] / fam_ms[k] / 1e6 / peak} for k in fam_ms}
    dom = max(fam_ms, key=fam_ms.get)
    total_frames = world * S * T * args.steps
    value = total_frames / (ms / 1e3)
    cpu_rate, cpu_t = cpu_reference(wl, args.ref_frames, os.cpu_count() or 1)[0] if world == 1 and not args.no_cpu else (None, None)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': args.workload, 'desc': wl['desc'], 'streams_per_gpu': S, 'frames_per_stream_per_step': T,
                   'frame': f'{W}x{H}x3 u8 BGR', 'window': win, 'rois': 2, 'windows': args.windows,
                   'window_jobs_per_step': jobs * world, 'l2_policy': 'inputs (51 GB of frames per GPU) larger than the 126 MB L2',
                   'parallelism': f'streams sharded x{world}, NCCL all-gather of per-stream records' if world > 1 else 'single GPU'},
        'windows_per_s': world * jobs * args.steps / (ms / 1e3),
        # F1 is the HBM-bound kernel the metric is quoted on ("ROI-sampled frames/s (%HBM peak)") and moves most of the
        # step's DRAM traffic; the dominant family BY TIME is FP64/issue-bound filter work (roofline_by_time).
        'roofline': {'kernel': 'roi (roi_staged_kernel)', 'bound': 'hbm', 'achieved': kernels['roi']['gbs'], 'peak': peak,
                     'unit': 'GB/s', 'frac': kernels['roi']['frac_hbm'], 'traffic': ROI_NCU_TRAFFIC_BYTES,
                     'peak_source': peak_src,
                     'note': 'achieved = algorithmic bytes (sum of 3*w*h over the ROIs of a launch + boxes + outputs) / mean '
                             'CUDA-event time of the F1 launches inside the timed region; traffic = dram__bytes_read+write of '
                             'one launch from ncu --set full (profiles/r1w_c2_summary.md, same boxes); ncu launch list: 61 us per launch'},
        'roofline_by_time': {'kernel': dom, 'bound': 'fp64 / issue (reported against HBM for reference)',
                             'achieved': kernels[dom]['gbs'], 'peak': peak, 'unit': 'GB/s', 'frac': kernels[dom]['frac_hbm'],
                             'note': 'dominant kernel family of the step by CUDA-event time'},
        'kernels': kernels,
        # libbpv kernels launched inside the timed region: the engine's step (roi, ring push, firls design, preprocess,
        # spectrum, xcorr) + pack_records for the result record
        'gpu_launches': (eng.launches_per_step + 1) * args.steps,
        'clocks': clk,
        'e2e': {'value': world * S * Te / e2e_s, 'unit': UNIT,
                'h2d_bytes_per_step': int(e2e_roi_bytes + host_boxes.numel() * 4 + host_ts.numel() * 8),
                'd2h_bytes_per_step': int(host_out.numel() * 8), 'frames_per_stream_per_step': Te,
                'how': 'BatchedSignalProcessor.roi_samples + step_signals on pinned HOST frames/boxes/timestamps; F1 reads the '
                       'ROI rows zero-copy over PCIe (the other 99.5 % of each frame never crosses the bus) on a side stream, '
                       'overlapped with the previous batch\'s window pipeline; records copied back to host every batch; wall clock'},
    }
    if cpu_rate is not None:
        line['cpu_baseline'] = {'value': cpu_rate, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                                'sample': f'{os.cpu_count()} streams x {args.ref_frames} steady-state frames (window prefilled), '
                                          f'oracle port of SignalProcessor.process per frame, one process per core; '
                                          f'{cpu_t:.1f} s of CPU work per core'}
    print(json.dumps(line), flush=True)


def _shutdown():
    try:
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized():
            tdist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='bpv', choices=['bpv', 'reference'])
    ap.add_argument('--workload', default='c2', choices=list(WORKLOADS))
    ap.add_argument('--windows', default='every_frame', choices=['every_frame', 'last'])
    ap.add_argument('--ref-frames', type=int, default=2048, help='CPU arm: steady-state frames per stream/core')
    ap.add_argument('--e2e-frames', type=int, default=8, help='frames per stream per e2e step (pinned host memory)')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        try:
            run_gpu(args, wl)
        finally:
            _shutdown()


if __name__ == '__main__':
    main()
