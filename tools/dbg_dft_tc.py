"""Dev check of the tcgen05 DFT-256 building block against numpy (float64)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
from bpv import ops
rng = np.random.default_rng(0)
for rows in (128, 77, 1000):
    z = (rng.standard_normal((rows, 256)) * np.hanning(256)[None, :]).astype(np.float32)
    d = ops.dft256_tc(torch.from_numpy(z).cuda()).cpu().numpy().astype(np.float64)
    torch.cuda.synchronize()
    X = np.fft.rfft(z.astype(np.float64), axis=1)
    ref = np.concatenate([X.real, -X.imag[:, 1:128]], axis=1)
    err = np.abs(d - ref).max(axis=1) / np.abs(ref).max(axis=1)
    print(rows, 'max rel-to-rowmax err', err.max(), 'median', np.median(err))
    if err.max() > 1e-3:
        bad = np.argwhere(np.abs(d - ref) > 1e-3 * np.abs(ref).max())
        print('bad entries', len(bad), bad[:10].tolist())
        print(d[0, :8], ref[0, :8])
