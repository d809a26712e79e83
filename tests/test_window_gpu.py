"""GPU parity of the window pipeline (F2 preprocessing, filter design, F3 spectra, F4 peaks / xcorr)
through the C-ABI, against the oracle on the same seeded windows.

Tolerances: north_star asks rtol 1e-4 for filtered signals / PSDs; the float64 kernels are held to a
much tighter bound here so that regressions show up.  Peak bins and lags must be bit-exact.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import bpv_oracle as orc
from tests import helpers as h

pytestmark = pytest.mark.gpu
RTOL = 1e-4          # the spec tolerance (north_star)
TIGHT = 1e-7         # what the float64 kernels actually achieve (regression guard)


def make_windows(seed, S, W, R=2, fps=30.0, irregular=True, fill=None, p_nan=0.03):
    """S streams, window W: timestamps [S,W] and raw samples [S,R,W] with NaN prefix (warm-up) of
    different lengths and missing detections."""
    from bpv import synth
    rng = np.random.default_rng(seed)
    t = np.full((S, W), np.nan)
    y = np.full((S, R, W), np.nan)
    for s in range(S):
        cnt = W if fill is None else fill[s % len(fill)]
        if cnt == 0:
            continue
        ts = synth.timestamps(rng, cnt, fps, irregular=irregular, drop=0.05 if irregular else 0.0, origin=rng.uniform(0, 50))
        t[s, W - cnt:] = ts
        y[s, :, W - cnt:] = synth.raw_signals(rng, ts, R=R, p_nan=p_nan if cnt > 8 else 0.0)
    return t, y


def to_ring(t, y):
    """Window arrays -> ring buffers with cap = W, head0 = W-1 (slot g % cap = g)."""
    return torch.from_numpy(t).cuda().contiguous(), torch.from_numpy(y).cuda().contiguous()


def params(S, R, W, methods, transform=orc.PGRAM_LS, **kw):
    from bpv import ops
    return ops.make_params(S, R, W, W, W - 1, 1, 1, methods, transform, **kw)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('order', [16, 8, 5, 4, 3, 2, 1])
def test_butter_design(order):
    from bpv import ops
    import scipy.signal
    fs = np.array([8.5, 9.7, 12.0, 25.3, 29.97, 30.0, 60.0, 120.0, 240.0, 5.0, 2.0])
    p = params(1, 1, 8, [], butter_order=order)
    got = ops.butter_sos_design(torch.from_numpy(fs).cuda(), p).cpu().numpy()
    for i, f in enumerate(fs):
        ref = orc.make_filter(orc.FILTER_BUTTER, f, butter_order=order)
        assert got[i].shape == ref.shape
        np.testing.assert_allclose(got[i], ref, rtol=1e-9, atol=1e-300, err_msg=f'fs={f} order={order}')


def test_firls_design():
    from bpv import ops
    fs = np.array([8.7, 10.0, 15.0, 25.3, 29.97, 30.0, 60.0, 120.0, 240.0])
    p = params(1, 1, 8, [])
    got = ops.firls_design(torch.from_numpy(fs).cuda(), p).cpu().numpy()
    for i, f in enumerate(fs):
        ref = orc.make_filter(orc.FILTER_FIR, f)
        np.testing.assert_allclose(got[i], ref, rtol=0, atol=1e-9 * np.abs(ref).max(), err_msg=f'fs={f}')
    # band edges the reference's firls rejects (ValueError) -> NaN taps
    bad = ops.firls_design(torch.tensor([7.5, 4.0], dtype=torch.float64, device='cuda'), p).cpu().numpy()
    assert np.isnan(bad).all()
    p2 = params(1, 1, 8, [], fir_taps=31)
    got = ops.firls_design(torch.from_numpy(fs).cuda(), p2).cpu().numpy()
    for i, f in enumerate(fs):
        ref = orc.make_filter(orc.FILTER_FIR, f, fir_taps=31)
        np.testing.assert_allclose(got[i], ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_firls_design_many_jobs_deterministic():
    """Several designs share a warp and a CTA's shared memory: 4099 designs (ragged last CTA) of cycling sampling rates
    must all equal the design of the same rate computed alone, bit for bit, on every run (caught a cross-warp
    shared-memory overlap once)."""
    from bpv import ops
    p = params(1, 1, 8, [])
    base = np.array([29.97, 8.7, 120.0, 30.0, 25.3, 61.7, 15.0])
    alone = torch.stack([ops.firls_design(torch.tensor([f], dtype=torch.float64, device='cuda'), p)[0] for f in base])
    fs = np.tile(base, 586)[:4099]
    for _ in range(3):
        got = ops.firls_design(torch.from_numpy(fs).cuda(), p)
        assert torch.equal(got, alone[torch.arange(4099, device='cuda') % len(base)])
    ref = orc.make_filter(orc.FILTER_FIR, 29.97)
    np.testing.assert_allclose(alone[0].cpu().numpy(), ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_dft256_tensor_core_block_matches_numpy():
    """bpv_dft256_tc: 256-point DFT of real segments as one tcgen05 contraction (3xTF32 split operands, fp32 TMEM
    accumulators) against numpy.fft.rfft in float64: error below 1e-5 of each row's maximum."""
    from bpv import ops
    rng = np.random.default_rng(3)
    for rows in (1, 77, 128, 1000):
        z = (rng.standard_normal((rows, 256)) * np.hanning(256)[None, :] * rng.uniform(1e-3, 1e3, (rows, 1))).astype(np.float32)
        d = ops.dft256_tc(torch.from_numpy(z).cuda()).cpu().numpy().astype(np.float64)
        X = np.fft.rfft(z.astype(np.float64), axis=1)
        ref = np.concatenate([X.real, -X.imag[:, 1:128]], axis=1)
        err = np.abs(d - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert err.max() < 1e-5, (rows, err.max())


METHOD_SETS = [
    [], [orc.DIFF_1], [orc.DIFF_2], [orc.DETREND_CONST], [orc.DETREND_LINEAR], [orc.INTERP_LINEAR], [orc.INTERP_CUBIC],
    [orc.FILTER_BUTTER], [orc.FILTER_FIR],
    [orc.DETREND_LINEAR, orc.FILTER_FIR],             # BASELINE config 2
    [orc.INTERP_CUBIC, orc.FILTER_BUTTER],            # BASELINE config 4
    [orc.INTERP_LINEAR, orc.DETREND_CONST, orc.FILTER_FIR],
    [orc.DIFF_1, orc.INTERP_CUBIC, orc.DETREND_LINEAR, orc.FILTER_BUTTER, orc.DIFF_2],
    [orc.FILTER_BUTTER, orc.INTERP_LINEAR],
]


@pytest.mark.parametrize('W,fps', [(64, 30.0), (250, 30.0), (300, 30.0), (500, 120.0)])   # 250 / 300: the specialised FIR tiles
@pytest.mark.parametrize('methods', METHOD_SETS, ids=lambda m: '-'.join(map(str, m)) or 'none')
def test_preprocess_matches_oracle(methods, W, fps):
    from bpv import ops
    S, R = 12, 2
    fill = [W, W, W - 1, W // 2, 130, 100, 5, 4, 3, 2, 1, 0]
    fill = [min(f, W) for f in fill]
    t, y = make_windows(W * 7 + len(methods), S, W, R, fps=fps, fill=fill)
    rt, ry = to_ring(t, y)
    p = params(S, R, W, methods)
    px, py, st = ops.window_preprocess(rt, ry, p)
    torch.cuda.synchronize()
    px, py, st = px.cpu().numpy(), py.cpu().numpy(), st.cpu().numpy()
    for s in range(S):
        for r in range(R):
            ex, ey = orc.preprocess(t[s], y[s, r], methods)
            assert h.close(px[s, r], ex, rtol=1e-12, atol_frac=0), (s, r, 'x')
            atol = 1e-11 * np.nanmax(np.abs(y[s, r])) if np.isfinite(y[s, r]).any() else 0.0   # eps-level of the raw DC
            assert h.close(py[s, r], ey, rtol=TIGHT, atol_frac=TIGHT, atol=atol), (s, r, fill[s], np.nanmax(np.abs(py[s, r] - ey)))
            n_valid = np.isfinite(y[s, r]).sum()
            assert st[s, r] == (0 if (n_valid >= 2 and np.isfinite(orc.window_fs(t[s]))) else 1)


def test_preprocess_sliding_jobs():
    """Several window jobs per stream reading one ring at consecutive heads (every-frame evaluation)."""
    from bpv import ops
    S, R, W, T = 3, 2, 48, 9
    cap = W + T
    rng = np.random.default_rng(3)
    from bpv import synth
    n_total = 70
    ts = np.stack([synth.timestamps(rng, n_total, 30.0, irregular=True, drop=0.05, origin=5.0) for _ in range(S)])
    ys = np.stack([synth.raw_signals(rng, ts[s], R=R, p_nan=0.05) for s in range(S)])
    ring_t = torch.full((S, cap), float('nan'), dtype=torch.float64, device='cuda')
    ring_y = torch.full((S, R, cap), float('nan'), dtype=torch.float64, device='cuda')
    methods = [orc.DETREND_LINEAR, orc.FILTER_BUTTER]
    g0 = 0
    while g0 < n_total:
        Tn = min(T, n_total - g0)
        ops.ring_push(ring_t, ring_y, g0, torch.from_numpy(ts[:, g0:g0 + Tn].copy()).cuda(),
                      torch.from_numpy(np.ascontiguousarray(ys[:, :, g0:g0 + Tn].transpose(0, 2, 1))).cuda())
        p = ops.make_params(S, R, cap, W, g0, 1, Tn, methods, orc.PGRAM_LS)
        px, py, st = ops.window_preprocess(ring_t, ring_y, p)
        px, py = px.cpu().numpy().reshape(S, Tn, R, W), py.cpu().numpy().reshape(S, Tn, R, W)
        for s in range(S):
            for j in range(Tn):
                head = g0 + j
                lo = head - W + 1
                tw = np.full(W, np.nan)
                tw[max(0, -lo):] = ts[s, max(lo, 0):head + 1]
                for r in range(R):
                    yw = np.full(W, np.nan)
                    yw[max(0, -lo):] = ys[s, r, max(lo, 0):head + 1]
                    ex, ey = orc.preprocess(tw, yw, methods)
                    assert h.close(px[s, j, r], ex, rtol=1e-12, atol_frac=0)
                    assert h.close(py[s, j, r], ey, rtol=TIGHT, atol_frac=TIGHT, atol=2e-9), (s, j, r)
        g0 += Tn


def test_cubic_duplicate_timestamps_flagged():
    """The reference raises ValueError (CubicSpline: x must be strictly increasing); we flag status 2."""
    from bpv import ops
    W = 32
    t, y = make_windows(11, 2, W, 2, irregular=False, p_nan=0.0)
    t[1, 10] = t[1, 9]
    rt, ry = to_ring(t, y)
    px, py, st = ops.window_preprocess(rt, ry, params(2, 2, W, [orc.INTERP_CUBIC]))
    st = st.cpu().numpy()
    assert (st[0] == 0).all() and (st[1] == 2).all()
    assert np.isnan(py.cpu().numpy()[1]).all()
    with pytest.raises(ValueError):
        orc.preprocess(t[1], y[1, 0], [orc.INTERP_CUBIC])


def test_unknown_method_raises():
    from bpv import ops
    t, y = make_windows(1, 1, 16, 1)
    rt, ry = to_ring(t, y)
    with pytest.raises(NotImplementedError):
        ops.window_preprocess(rt, ry, params(1, 1, 16, [99]))


# ------------------------------------------------------------------------------------------------
# F3 / F4: kernels fed with the ORACLE's processed windows, so each kernel is checked in isolation
# ------------------------------------------------------------------------------------------------
def oracle_proc(t, y, methods):
    S, R, W = y.shape
    px, py = np.empty((S, R, W)), np.empty((S, R, W))
    for s in range(S):
        for r in range(R):
            px[s, r], py[s, r] = orc.preprocess(t[s], y[s, r], methods)
    return px, py


SPEC_CASES = [
    (orc.DFT_RFFT, [orc.DIFF_1], {}), (orc.DFT_RFFT, [orc.FILTER_BUTTER], {}), (orc.DFT_RFFT, [], {}),
    (orc.PGRAM_WELCH, [orc.DETREND_LINEAR, orc.FILTER_FIR], {}), (orc.PGRAM_WELCH, [], {}),
    (orc.PGRAM_LS, [orc.FILTER_BUTTER], dict(min_freq=0.7)), (orc.PGRAM_LS, [], {}),
    (orc.PGRAM_LS, [orc.INTERP_CUBIC, orc.FILTER_BUTTER], {}), (orc.PGRAM_LS, [orc.DETREND_CONST], dict(ls_num_freqs=512)),
]


@pytest.mark.parametrize('W,fps', [(64, 30.0), (300, 30.0), (600, 120.0)])
@pytest.mark.parametrize('transform,methods,kw', SPEC_CASES, ids=lambda v: str(v).replace(' ', ''))
def test_spectrum_and_peak_match_oracle(transform, methods, kw, W, fps):
    from bpv import ops
    S, R = 14, 2
    fill = [min(f, W) for f in [W, W, W - 1, W // 2, 257, 256, 255, 130, 40, 5, 4, 3, 2, 1]]
    t, y = make_windows(W * 13 + transform, S, W, R, fps=fps, fill=fill)
    px, py = oracle_proc(t, y, methods)
    p = params(S, R, W, methods, transform, **kw)
    dx, dy = torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda()
    for store in (True, False):
        o = ops.window_spectrum(dx, dy, p, store=store)
        torch.cuda.synchronize()
        nb, pi = o['num_bins'].cpu().numpy(), o['peak_idx'].cpu().numpy()
        pf, pm = o['peak_freq'].cpu().numpy(), o['peak_mag'].cpu().numpy()
        for s in range(S):
            for r in range(R):
                ef, em = orc.spectrum(px[s, r], py[s, r], transform, **kw)
                ex, ey, ei = orc.peak(ef, em)
                assert nb[s, r] == len(ef), (s, r)
                n_valid = int(np.isfinite(py[s, r]).sum())
                if store:
                    gf, gm = o['freqs'].cpu().numpy()[s, r, :len(ef)], o['mags'].cpu().numpy()[s, r, :len(ef)]
                    assert h.close(gf, ef, rtol=1e-6, atol_frac=0), (s, r, 'freqs')
                    if not (transform == orc.PGRAM_LS and n_valid < 4):
                        assert h.close(gm, em, rtol=RTOL, atol_frac=1e-5, atol=1e-30), (s, r, fill[s], 'mags', np.nanmax(np.abs(gm - em)))
                if transform == orc.PGRAM_LS and n_valid < 4:
                    continue   # n <= 3: the floating-mean LS model fits exactly, p == 1 up to rounding at every bin
                assert pi[s, r] == ei, (s, r, fill[s], pi[s, r], ei)
                assert h.same(np.isnan(pf[s, r]), np.isnan(ex))
                if ei >= 0:
                    # LS grid frequencies are exact; DFT/Welch bins scale with fs = 1/mean(diff(t)), where numpy's
                    # pairwise mean and our (m-1)/(t_last-t_first) may differ in the last ulp
                    np.testing.assert_allclose(pf[s, r], ex, rtol=0 if transform == orc.PGRAM_LS else 4e-16 * 8)
                    np.testing.assert_allclose(pm[s, r], ey, rtol=1e-7, atol=1e-9 * max(1.0, abs(ey)))


@pytest.mark.parametrize('W,fps', [(40, 30.0), (300, 30.0), (1200, 120.0)])
@pytest.mark.parametrize('methods,force', [([orc.INTERP_CUBIC, orc.FILTER_BUTTER], None), ([orc.INTERP_LINEAR, orc.DETREND_CONST], None),
                                           ([orc.DIFF_1], '1'), ([], '1')], ids=['cubic_butter', 'linear_const', 'diff1_forced', 'none_forced'])
def test_dft_tensor_core_path_matches_oracle(methods, force, W, fps, monkeypatch):
    """DFT_RFFT as a tcgen05 contraction (dft_tc_kernel: 3xTF32, one twiddle matrix per launch because full windows
    share n = W) + float64 decision (dft_peak_kernel); default for resampling pipelines, forced here for the others.
    Peak bins exact, magnitudes within the spec tolerance, short / holed windows through the float64 kernel."""
    from bpv import ops
    if force is not None:
        monkeypatch.setenv('BPV_DFT_TC', force)
    S, R = 12, 2
    fill = [min(f, W) for f in [W, W, W, W - 1, W // 2, 257, 130, 40, 5, 3, 2, 1]]
    t, y = make_windows(W * 19 + len(methods), S, W, R, fps=fps, fill=fill, p_nan=0.0 if not methods else 0.03)
    if not methods:
        y[1, 0, W // 3] = np.nan                        # a hole in an otherwise full window -> float64 kernel
    px, py = oracle_proc(t, y, methods)
    p = params(S, R, W, methods, orc.DFT_RFFT)
    dx, dy = torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda()
    for store in (True, False):
        o = ops.window_spectrum(dx, dy, p, store=store)
        torch.cuda.synchronize()
        nb, pi = o['num_bins'].cpu().numpy(), o['peak_idx'].cpu().numpy()
        pf, pm = o['peak_freq'].cpu().numpy(), o['peak_mag'].cpu().numpy()
        for s in range(S):
            for r in range(R):
                ef, em = orc.spectrum(px[s, r], py[s, r], orc.DFT_RFFT)
                ex, ey, ei = orc.peak(ef, em)
                assert nb[s, r] == len(ef), (s, r, nb[s, r], len(ef))
                if store:
                    gf, gm = o['freqs'].cpu().numpy()[s, r, :len(ef)], o['mags'].cpu().numpy()[s, r, :len(ef)]
                    assert h.close(gf, ef, rtol=1e-6, atol_frac=0), (s, r, 'freqs')
                    assert h.close(gm, em, rtol=RTOL, atol_frac=1e-5, atol=1e-30), (s, r, fill[s], 'mags', np.nanmax(np.abs(gm - em)))
                assert pi[s, r] == ei, (s, r, fill[s], pi[s, r], ei)
                if ei >= 0:
                    np.testing.assert_allclose(pf[s, r], ex, rtol=4e-16 * 8)
                    np.testing.assert_allclose(pm[s, r], ey, rtol=1e-7, atol=1e-9 * max(1.0, abs(ey)))


@pytest.mark.parametrize('W,fps', [(64, 30.0), (250, 30.0), (300, 30.0), (600, 120.0)])   # 250 / 300: specialised xcorr leading dimensions
@pytest.mark.parametrize('methods', [[orc.FILTER_BUTTER], [orc.DETREND_LINEAR, orc.FILTER_FIR], [orc.INTERP_CUBIC, orc.FILTER_BUTTER], []],
                         ids=lambda m: '-'.join(map(str, m)) or 'none')
def test_xcorr_matches_oracle(methods, W, fps):
    from bpv import ops
    S, R = 10, 3
    fill = [min(f, W) for f in [W, W - 1, W // 2, 130, 40, 5, 3, 2, 1, 0]]
    t, y = make_windows(W * 17 + len(methods), S, W, R, fps=fps, fill=fill)
    px, py = oracle_proc(t, y, methods)
    p = params(S, R, W, methods)
    dx, dy = torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda()
    pairs = list(itertools.combinations(range(R), 2))
    for store in (True, False):
        o = ops.window_xcorr(dx, dy, p, store=store)
        torch.cuda.synchronize()
        nl, li = o['num_lags'].cpu().numpy(), o['lag_idx'].cpu().numpy()
        ls, lc = o['lag_sec'].cpu().numpy(), o['lag_corr'].cpu().numpy()
        for s in range(S):
            for k, (a, b) in enumerate(pairs):
                el, ec = orc.xcorr(px[s, a], py[s, a], py[s, b])
                ex, ey, ei = orc.peak(el, ec)
                assert nl[s, k] == len(el)
                if store:
                    assert h.close(o['lags'].cpu().numpy()[s, k, :len(el)], el, rtol=1e-6, atol_frac=1e-7)
                    assert h.close(o['corr'].cpu().numpy()[s, k, :len(el)], ec, rtol=RTOL, atol_frac=1e-5)
                assert li[s, k] == ei, (s, k, li[s, k], ei)
                if ei >= 0:
                    assert ls[s, k] == ex
                    np.testing.assert_allclose(lc[s, k], ey, rtol=1e-9)
                else:
                    assert np.isnan(ls[s, k]) and np.isnan(lc[s, k])


def test_unknown_transform_raises():
    from bpv import ops
    x = torch.zeros((1, 1, 16), dtype=torch.float64, device='cuda')
    with pytest.raises(NotImplementedError):
        ops.window_spectrum(x, x, params(1, 1, 16, [], 9))
