"""cProfile of the drop-in SignalProcessor.process() on ONE 640x480 stream (the shape of bench.py's latency_c1):
where the per-frame host time goes.  Usage: python tools/profile_process.py [frames]"""
import cProfile, io, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
import torch
import signal_processor as sp
from bpv import synth

frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(3)
H, W, win = 480, 640, 300
proc = sp.SignalProcessor(None, 1, win, 50, color_channel=sp.SignalColorChannel.GREEN,
                          processing_methods=[sp.SignalProcessingMethod.FILTER_BUTTER],
                          spectrum_transform=sp.SignalSpectrumTransform.PGRAM_LS, min_freq=0.7)
n = win + frames_n + 8
ts = synth.timestamps(rng, n, 30.0)
det = synth.detections(rng, n, H, W)
pool = synth.frames(rng, ts[:8], H, W)
for i in range(win + 8):
    proc.process(synth.FrameData(pool[i % 8], float(ts[i])), synth.ModelResults(det, i))
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(win + 8, win + 8 + frames_n // 2):
    proc.process(synth.FrameData(pool[i % 8], float(ts[i])), synth.ModelResults(det, i))
print('ms per frame (no profiler): %.3f' % ((time.perf_counter() - t0) / (frames_n // 2) * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(win + 8 + frames_n // 2, n):
    proc.process(synth.FrameData(pool[i % 8], float(ts[i])), synth.ModelResults(det, i))
pr.disable()
for key in ('cumulative', 'tottime'):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6000])
