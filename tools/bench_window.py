"""Development micro-benchmark of the window pipeline (F2/F3/F4) on signals-only input (no frames)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
from bpv import _cabi, ops, synth  # noqa: E402
from bpv.engine import BatchedSignalProcessor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--S', type=int, default=256)
ap.add_argument('--T', type=int, default=32)
ap.add_argument('--W', type=int, default=300)
ap.add_argument('--fps', type=float, default=30.0)
ap.add_argument('--methods', default='DETREND_LINEAR,FILTER_FIR')
ap.add_argument('--transform', default='PGRAM_WELCH')
ap.add_argument('--ls-num-freqs', type=int, default=0)
ap.add_argument('--irregular', action='store_true')
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--windows', default='every_frame')
a = ap.parse_args()
methods = [getattr(_cabi, m) for m in a.methods.split(',') if m]
eng = BatchedSignalProcessor(a.S, 2, signal_max_samples=a.W, max_frames_per_step=a.T, processing_methods=methods,
                             spectrum_transform=getattr(_cabi, a.transform), ls_num_freqs=a.ls_num_freqs or None, windows=a.windows)
rng = np.random.default_rng(0)
n = a.W + a.T * (a.iters + 5)
ts = np.stack([synth.timestamps(rng, n, a.fps, irregular=a.irregular, drop=0.05 if a.irregular else 0) for _ in range(a.S)])
ys = np.stack([synth.raw_signals(rng, ts[s]).T for s in range(a.S)])
ts_d, ys_d = torch.from_numpy(ts).cuda(), torch.from_numpy(ys).cuda()
eng.windows, keep = 'last', eng.windows
g = 0
while g < a.W:
    eng.step_signals(ys_d[:, g:g + a.T].contiguous(), ts_d[:, g:g + a.T].contiguous())
    g += a.T
eng.windows = keep
fam = {}
orig = {k: getattr(ops, k) for k in ('ring_push', 'window_design', 'window_filter', 'window_spectrum', 'window_xcorr')}


def timed(fn, key):
    def wrap(*aa, **kk):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*aa, **kk)
        e1.record()
        fam.setdefault(key, []).append((e0, e1))
        return out
    return wrap


for k, fn in orig.items():
    setattr(ops, k, timed(fn, k))
for i in range(a.iters + 3):
    if i == 3:
        fam.clear()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
    eng.step_signals(ys_d[:, g:g + a.T].contiguous(), ts_d[:, g:g + a.T].contiguous())
    g += a.T
s1.record()
torch.cuda.synchronize()
tot = s0.elapsed_time(s1) / a.iters
jobs = a.S * (a.T if a.windows == 'every_frame' else 1)
print(f'S={a.S} T={a.T} W={a.W} methods={a.methods} transform={a.transform}: {tot:.3f} ms/step, {jobs / tot * 1e3:.0f} windows/s')
for k, v in fam.items():
    print(f'  {k:20s} {np.mean([x.elapsed_time(y) for x, y in v]):.4f} ms')
