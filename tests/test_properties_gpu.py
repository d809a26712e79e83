"""GPU: size-independent properties at BASELINE.json's full sizes (where running the oracle on everything would
take hours): integer checksums for F1, linearity / shift invariance / symmetry for the window pipeline, and the
equivalence of the two evaluation modes of the engine.  A sample of the units is still compared with the oracle."""
import numpy as np
import pytest
import torch

from oracle import bpv_oracle as orc
from tests import helpers as h

pytestmark = pytest.mark.gpu


def test_roi_checksums_full_hd_batch():
    """Config-2 frame geometry, 512 frames: sums over a partition of the frame add up to the whole-frame sums
    (torch integer sum), N is the box area, and a sample of ROIs matches the oracle bit for bit."""
    from bpv import ops, synth
    N, H, W = 512, 1080, 1920
    g = torch.Generator(device='cuda'); g.manual_seed(3)
    frames = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device='cuda', generator=g)
    # 4 ROIs that tile the frame exactly (unaligned split points), + the whole frame
    xs, ys = 613, 407
    tiles = [(0, 0, xs, ys), (xs, 0, W, ys), (0, ys, xs, H), (xs, ys, W, H), (0, 0, W, H)]
    boxes = torch.tensor(tiles, dtype=torch.int32, device='cuda')[None].repeat(N, 1, 1).contiguous()
    val, sums = ops.roi_sample(frames, boxes, orc.CHROM_GREEN, want_sums=True, roi_pixels_hint=H * W // 4)
    torch.cuda.synchronize()
    whole = frames.view(N, H * W, 3).sum(dim=1, dtype=torch.int64)            # [N, 3] B, G, R
    assert torch.equal(sums[:, 4, :3], whole)
    assert torch.equal(sums[:, :4, :3].sum(dim=1), whole)
    assert torch.equal(sums[:, :, 3], torch.tensor([xs * ys, (W - xs) * ys, xs * (H - ys), (W - xs) * (H - ys), W * H],
                                                   device='cuda')[None].expand(N, 5))
    # value == (2G - B - R + 2N) / 4N computed in float64 from the integer sums
    s = sums.double()
    exp = (2 * s[..., 1] - s[..., 0] - s[..., 2] + 2 * s[..., 3]) / (4 * s[..., 3])
    assert torch.equal(val, exp)
    # synthetic forehead/palm boxes: a sample against the oracle
    rng = np.random.default_rng(0)
    bx = synth.roi_boxes(rng, N, H, W, p_none=0.05, p_oob=0.05)
    v2, s2 = ops.roi_sample(frames, torch.from_numpy(bx).cuda(), orc.GREEN, want_sums=True, roi_pixels_hint=6000)
    for f in rng.choice(N, 12, replace=False):
        fr = frames[f].cpu().numpy()
        for r in range(2):
            if bx[f, r, 0] == synth.NO_BOX:
                assert torch.isnan(v2[f, r])
                continue
            assert tuple(int(v) for v in s2[f, r].cpu()) == orc.roi_sums(fr, bx[f, r])


def _engine(S, W, methods, transform, T=1, **kw):
    from bpv.engine import BatchedSignalProcessor
    return BatchedSignalProcessor(S, 2, signal_max_samples=W, max_frames_per_step=max(T, W), processing_methods=methods,
                                  spectrum_transform=transform, windows='last', store_arrays=True, **kw)


def _run(eng, ts, ys):
    """ts [S, n], ys [S, n, R] device tensors, n <= W pushed in one step; returns the StepResult arrays."""
    eng.reset()
    return eng.step_signals(ys.contiguous(), ts.contiguous())


@pytest.mark.parametrize('methods,transform', [
    ([orc.DETREND_LINEAR, orc.FILTER_FIR], orc.PGRAM_WELCH),     # config 2
    ([orc.FILTER_BUTTER], orc.PGRAM_LS),                         # config 1 / 5
    ([orc.INTERP_CUBIC, orc.FILTER_BUTTER], orc.DFT_RFFT),
], ids=['c2', 'c5', 'cubic_butter_dft'])
def test_window_pipeline_properties_many_streams(methods, transform):
    """4096 streams x W=300 (8192 signals): linearity of the (linear) preprocessing, invariances of the spectra and
    of the xcorr normalisation, and an oracle spot check."""
    from bpv import synth
    S, W, R = 4096, 300, 2
    rng = np.random.default_rng(11)
    base_t = synth.timestamps(rng, W, 30.0, irregular=True, drop=0.03)
    jitter = rng.uniform(-2e-3, 2e-3, (S, W))
    ts = np.sort(base_t[None, :] + jitter, axis=1) + rng.uniform(0, 500, (S, 1))
    ys = 120 + rng.standard_normal((S, W, R)).cumsum(axis=1) * 0.05 + 0.5 * np.sin(2 * np.pi * 1.3 * (ts - ts[:, :1]))[..., None]
    ys[rng.uniform(size=ys.shape) < 0.01] = np.nan
    t_d, y_d = torch.from_numpy(ts).cuda(), torch.from_numpy(ys).cuda()
    eng = _engine(S, W, methods, transform)
    a = _run(eng, t_d, y_d)
    py = a.arrays['proc_y'].clone()
    mags, pidx, lidx = a.arrays['mags'].clone(), a.peak_idx.clone(), a.lag_idx.clone()
    corr = a.arrays['corr'].clone()
    assert int((a.status != 0).sum()) == 0
    # (1) linearity: every method here is linear in y => P(alpha*y) == alpha*P(y); normalised outputs unchanged
    alpha = 3.0
    b = _run(eng, t_d, y_d * alpha)
    scale = float(py[torch.isfinite(py)].abs().max())
    assert torch.allclose(b.arrays['proc_y'], alpha * py, rtol=1e-9, atol=1e-9 * scale, equal_nan=True)
    assert torch.equal(b.lag_idx, lidx)                                   # xcorr is scale free
    nl = a.arrays['num_lags'].clone()
    lvalid = torch.arange(corr.shape[-1], device=corr.device)[None, None, :] < nl[..., None]      # entries past num_lags are undefined
    assert torch.allclose(torch.where(lvalid, b.arrays['corr'], 0), torch.where(lvalid, corr, 0), rtol=1e-5, atol=1e-6, equal_nan=True)
    if transform == orc.PGRAM_LS:
        assert torch.equal(b.peak_idx, pidx)                              # normalised LS is scale free
        nb = a.arrays['num_bins'].clone()
        valid = torch.arange(mags.shape[-1], device=mags.device)[None, None, :] < nb[..., None]    # entries past num_bins are undefined
        assert torch.equal(b.arrays['num_bins'], nb)
        assert torch.allclose(torch.where(valid, b.arrays['mags'], 0), torch.where(valid, mags, 0), rtol=1e-4, atol=2e-5, equal_nan=True)
    # (2) time-shift invariance: shifting all timestamps by a constant changes nothing but proc_x
    c = _run(eng, t_d + 1000.0, y_d)
    assert torch.allclose(c.arrays['proc_y'], py, rtol=1e-6, atol=1e-7 * scale, equal_nan=True)
    same = (c.peak_idx == pidx).double().mean().item()
    assert same > 0.999, same          # fs changes in the last bits with the shift; near-ties may flip at the 1e-3 level
    # (3) spot check against the oracle
    for s in rng.choice(S, 6, replace=False):
        for r in range(R):
            ex, ey = orc.preprocess(ts[s], ys[s, :, r], methods)
            assert h.close(py[s, r].cpu().numpy(), ey, rtol=1e-6, atol_frac=1e-7)
            ef, em = orc.spectrum(ex, ey, transform)
            assert int(pidx[s, r]) == orc.peak(ef, em)[2]


def test_xcorr_symmetry_many_pairs():
    """corr(a, b)[k] == corr(b, a)[-k] for 8192 pairs; the lag of the maximum mirrors."""
    from bpv import ops
    J, W = 8192, 300
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    y = torch.randn((J, 2, W), dtype=torch.float64, device='cuda', generator=g)
    y[:, 1, 3:] += 0.8 * y[:, 0, :-3]                      # b lags a by 3 samples
    x = (torch.arange(W, dtype=torch.float64, device='cuda') / 30.0)[None, None, :].expand(J, 2, W).contiguous()
    p = ops.make_params(J, 2, W, W, W - 1, 1, 1, [], orc.PGRAM_LS)
    o1 = ops.window_xcorr(x, y, p)
    c1, l1 = o1['corr'].clone(), o1['lag_idx'].clone()
    o2 = ops.window_xcorr(x, y.flip(1).contiguous(), p)
    assert torch.allclose(o2['corr'], c1.flip(-1), rtol=1e-6, atol=1e-7)
    assert torch.equal(o2['lag_idx'], (2 * W - 2) - l1)
    assert int((l1[:, 0] == (W - 1) - 3).sum()) > 0.99 * J         # the planted 3-sample lag wins


def test_every_frame_equals_last_mode():
    """Evaluating the window after every frame of a T-frame step == T single-frame steps (ring addressing)."""
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    S, W, T, R = 64, 300, 32, 2
    rng = np.random.default_rng(2)
    n = W + 2 * T
    ts = np.stack([synth.timestamps(rng, n, 30.0, irregular=True, drop=0.05) for _ in range(S)])
    ys = np.stack([synth.raw_signals(rng, ts[s], R=R, p_nan=0.02).T for s in range(S)])
    t_d, y_d = torch.from_numpy(ts).cuda(), torch.from_numpy(ys).cuda()
    kw = dict(signal_max_samples=W, processing_methods=[orc.DETREND_LINEAR, orc.FILTER_FIR], spectrum_transform=orc.PGRAM_WELCH)
    e1 = BatchedSignalProcessor(S, R, max_frames_per_step=T, windows='every_frame', **kw)
    e2 = BatchedSignalProcessor(S, R, max_frames_per_step=1, windows='last', **kw)
    for g0 in range(0, n - T + 1, T):
        r1 = e1.step_signals(y_d[:, g0:g0 + T].contiguous(), t_d[:, g0:g0 + T].contiguous())
        bpm1 = r1.bpm.view(S, T, R).clone(); ptt1 = r1.ptt_ms.view(S, T, 1).clone()
        for j in range(T):
            r2 = e2.step_signals(y_d[:, g0 + j:g0 + j + 1].contiguous(), t_d[:, g0 + j:g0 + j + 1].contiguous())
            assert torch.equal(torch.nan_to_num(bpm1[:, j], nan=-1.0), torch.nan_to_num(r2.bpm, nan=-1.0)), (g0, j)
            assert torch.equal(torch.nan_to_num(ptt1[:, j], nan=-1.0), torch.nan_to_num(r2.ptt_ms, nan=-1.0)), (g0, j)


def test_config5_scale_65536_streams_signals_only():
    """BASELINE configs[4] stream count on one GPU (signals only): 65 536 streams x 2 ROIs, Butterworth + LS, one
    window job per stream; every result finite, and a sample equal to the oracle."""
    from bpv import synth
    S, W, R = 65536, 300, 2
    rng = np.random.default_rng(4)
    t1 = synth.timestamps(rng, W, 30.0)
    f = rng.uniform(0.8, 3.0, (S, 1))
    ts = t1[None, :] + rng.uniform(0, 100, (S, 1))
    ys = (120 + 0.5 * np.sin(2 * np.pi * f * t1[None, :]) + 0.1 * rng.standard_normal((S, W)))[..., None] + np.zeros((1, 1, R))
    ys[..., 1] = np.roll(ys[..., 0], 1, axis=1) + 0.05 * rng.standard_normal((S, W))
    eng = _engine(S, W, [orc.FILTER_BUTTER], orc.PGRAM_LS, min_freq=0.7)
    eng.store_arrays = False
    res = _run(eng, torch.from_numpy(ts).cuda(), torch.from_numpy(ys).cuda())
    bpm, ptt = res.bpm.cpu().numpy(), res.ptt_ms.cpu().numpy()
    assert np.isfinite(bpm).all() and np.isfinite(ptt).all()
    grid = np.linspace(0.7, 4.0, W) * 60
    err = np.abs(bpm[:, 0] - f[:, 0] * 60)
    assert np.median(err) <= (grid[1] - grid[0])                   # the planted heart rate is recovered to a bin
    pidx = res.peak_idx.cpu().numpy()
    for s in rng.choice(S, 5, replace=False):
        for r in range(R):
            ex, ey = orc.preprocess(ts[s], ys[s, :, r], [orc.FILTER_BUTTER], min_freq=0.7)
            ef, em = orc.spectrum(ex, ey, orc.PGRAM_LS, min_freq=0.7)
            assert pidx[s, r] == orc.peak(ef, em)[2]
