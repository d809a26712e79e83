"""Micro-benchmark of the F1 kernel alone (development aid; bench.py is the judged harness)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
from bpv import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=2048)
ap.add_argument('--H', type=int, default=1080)
ap.add_argument('--W', type=int, default=1920)
ap.add_argument('--iters', type=int, default=20)
ap.add_argument('--hint', type=int, default=0)
ap.add_argument('--mode', type=int, default=1)
ap.add_argument('--l2gran', type=int, default=0)
ap.add_argument('--host', action='store_true', help='frames in pinned host memory (zero-copy reads over PCIe)')
ap.add_argument('--dirty', type=int, default=0, help='MB of an L2-resident scratch buffer rewritten (left dirty in L2) before every timed launch')
ap.add_argument('--box', default='', help='w,h: fixed-size boxes at random positions instead of the synthetic forehead/palm boxes')
a = ap.parse_args()
N, H, W = a.frames, a.H, a.W
from bpv import _cabi  # noqa: E402
torch.cuda.init(); torch.zeros(1, device='cuda')
if a.l2gran:
    _cabi.check(_cabi.lib().bpv_set_l2_fetch_granularity(a.l2gran), 'l2gran')
print('l2 fetch granularity', _cabi.lib().bpv_get_l2_fetch_granularity())
if a.host:
    frames = torch.empty((N, H, W, 3), dtype=torch.uint8, pin_memory=True)
    tmp = torch.empty((64, H, W, 3), dtype=torch.uint8, device='cuda')
    for i in range(0, N, 64):
        tmp.random_(0, 256)
        frames[i:i + 64].copy_(tmp[:min(64, N - i)])
    del tmp
else:
    frames = torch.empty((N, H, W, 3), dtype=torch.uint8, device='cuda')
    for i in range(0, N, 64):
        frames[i:i + 64].random_(0, 256)
rng = np.random.default_rng(0)
boxes_np = synth.roi_boxes(rng, N, H, W)
if a.box:
    bw, bh = (int(v) for v in a.box.split(','))
    x0 = rng.integers(0, W - bw + 1, (N, 2)); y0 = rng.integers(0, H - bh + 1, (N, 2))
    boxes_np = np.stack([x0, y0, x0 + bw, y0 + bh], axis=-1).astype(np.int32)
boxes = torch.from_numpy(boxes_np).cuda()


def _sl(a, b, n):
    a, b, _ = slice(int(a), int(b)).indices(n)
    return a, max(a, b)


nbytes = 0
for f in range(N):
    for r in range(2):
        b = boxes_np[f, r]
        if b[0] == synth.NO_BOX:
            continue
        xa, xb = _sl(b[0], b[2], W)
        ya, yb = _sl(b[1], b[3], H)
        nbytes += 3 * (xb - xa) * (yb - ya)
hint = a.hint or int(nbytes / 3 / (2 * N))
out = torch.empty((N, 2), dtype=torch.float64, device='cuda')
for _ in range(3):
    ops.roi_sample(frames, boxes, a.mode, roi_pixels_hint=hint, out_value=out)
torch.cuda.synchronize()
print('checksum', repr(float(out.nan_to_num().double().sum().item())), int(out.isnan().sum().item()))
if a.dirty:
    scratch = torch.zeros(a.dirty << 20, dtype=torch.uint8, device='cuda')
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters)]
    for i in range(a.iters):
        scratch.add_(1)                      # read-modify-write: the whole buffer is dirty in L2 when F1 starts
        e0[i].record()
        ops.roi_sample(frames, boxes, a.mode, roi_pixels_hint=hint, out_value=out)
        e1[i].record()
    torch.cuda.synchronize()
    ts = [e0[i].elapsed_time(e1[i]) for i in range(a.iters)]
else:
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    ev[0].record()
    for i in range(a.iters):
        ops.roi_sample(frames, boxes, a.mode, roi_pixels_hint=hint, out_value=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
t = float(np.median(ts))
print(f'frames={N} {W}x{H} hint={hint} roi_bytes={nbytes/1e6:.1f} MB  median {t*1e3:.1f} us  min {min(ts)*1e3:.1f} us '
      f'-> {nbytes/t/1e6:.1f} GB/s median, {nbytes/min(ts)/1e6:.1f} GB/s best, {N/t*1e3:.0f} frames/s')
