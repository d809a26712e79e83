import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
from oracle import bpv_oracle as orc
from bpv import ops
from tests.test_window_gpu import make_windows, to_ring, params
for methods, W, fps in [([], 500, 120.0), ([6], 64, 30.0), ([7], 64, 30.0), ([7], 300, 30.0), ([4, 7], 300, 30.0)]:
    S, R = 12, 2
    fill = [min(f, W) for f in [W, W, W - 1, W // 2, 130, 100, 5, 4, 3, 2, 1, 0]]
    t, y = make_windows(W * 7 + len(methods), S, W, R, fps=fps, fill=fill)
    rt, ry = to_ring(t, y)
    px, py, st = ops.window_preprocess(rt, ry, params(S, R, W, methods))
    px, py, st = px.cpu().numpy(), py.cpu().numpy(), st.cpu().numpy()
    print('methods', methods, W, fps)
    for s in range(S):
        for r in range(R):
            ex, ey = orc.preprocess(t[s], y[s, r], methods)
            nanmis = int((np.isnan(py[s, r]) != np.isnan(ey)).sum())
            ok = np.isfinite(ey) & np.isfinite(py[s, r])
            err = np.abs(py[s, r] - ey)[ok].max() if ok.any() else 0
            errx = np.nanmax(np.abs(px[s, r] - ex)) if np.isfinite(ex).any() else 0
            sc = np.abs(ey[ok]).max() if ok.any() else 0
            print(f'  s={s} r={r} fill={fill[s]} nvalid={np.isfinite(y[s,r]).sum()} st={st[s,r]} nanmis={nanmis} maxerr={err:.3e} scale={sc:.3e} errx={errx:.2e}')
