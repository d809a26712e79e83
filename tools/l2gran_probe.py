"""Dev probe: does cudaLimitMaxL2FetchGranularity / ld .L2::64B change DRAM traffic of the ROI kernel?"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
from bpv import _cabi
g = int(os.environ.get('L2G', '0'))
if g:
    print('set before context:', _cabi.lib().bpv_set_l2_fetch_granularity(g), _cabi.lib().bpv_get_l2_fetch_granularity())
import numpy as np, torch
from bpv import ops, synth
N, H, W = 4096, 1080, 1920
frames = torch.empty((N, H, W, 3), dtype=torch.uint8, device='cuda')
for i in range(0, N, 64):
    frames[i:i + 64].random_(0, 256)
print('granularity now', _cabi.lib().bpv_get_l2_fetch_granularity())
boxes = torch.from_numpy(synth.roi_boxes(np.random.default_rng(0), N, H, W)).cuda()
for _ in range(4):
    ops.roi_sample(frames, boxes, 1, roi_pixels_hint=5800)
torch.cuda.synchronize()
