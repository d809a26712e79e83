"""GPU: the batched engine end to end (F1 -> ring -> F2 -> F3 -> F4) against (a) the golden outputs of
the unmodified reference and (b) the oracle on multi-stream, multi-frame-per-step sequences."""
import numpy as np
import pytest
import torch

from oracle import bpv_oracle as orc
from tests import helpers as h

pytestmark = pytest.mark.gpu

# Windows with fewer than MIN_N valid samples are numerically degenerate IN THE REFERENCE: a 16th-order
# band-pass / 127-tap FIR of 2-3 samples leaves ~1e-13 of rounding residue, the floating-mean Lomb-Scargle
# model fits <= 3 points exactly (p == 1 at every bin) and the 3-5 xcorr lags are near-tied, so the argmax is
# decided by the last bits of scipy's summation order.  Values are still compared (with residue-level atol);
# bit-exact peak bins / lags are required from MIN_N samples on.
MIN_N = 4      # evidence: tests/test_degenerate_windows_cpu.py (the reference flips its own argmax under a 1-ulp input change below 4 samples, never from 4 on)
# ... and a processed window whose amplitude is below RESIDUE x the raw level is rounding noise of the
# filter (e.g. order-16 Butterworth at 120 fps on < ~20 samples: output ~1e-22 for an input of 140).
RESIDUE = 1e-9


def _box4(b6):
    from bpv import synth
    b6 = np.asarray(b6, dtype=float)
    if np.isnan(b6).any():
        return [synth.NO_BOX, 0, 0, 0]
    return [int(v) for v in b6[2:6]]


@pytest.mark.parametrize('name', list(h.CASES))
def test_golden_replay(name):
    """One stream, one frame per step: every per-frame output the reference produced."""
    from bpv.engine import BatchedSignalProcessor
    channel, methods, transform, window, n, fps, irregular, p_none, roi_ms, kw = h.CASES[name]
    g = h.load_case(name)
    frames = h.case_frames(g)
    eng = BatchedSignalProcessor(1, 2, signal_max_samples=window, max_frames_per_step=1, color_channel=h.CHANNEL[channel],
                                 processing_methods=[h.METHOD[m] for m in methods], spectrum_transform=h.TRANSFORM[transform],
                                 store_arrays=True, **kw)
    full_at = set(int(i) for i in g['full_at'])
    ls = transform == 'PGRAM_LS'
    for i in range(n):
        boxes = torch.tensor([[[_box4(b) for b in g['boxes'][i]]]], dtype=torch.int32, device='cuda')
        fr = torch.from_numpy(frames[i:i + 1][None]).cuda()
        ts = torch.tensor([[g['ts'][i]]], dtype=torch.float64, device='cuda')
        res = eng.step(fr, boxes, ts)
        assert h.same(res.samples.cpu().numpy()[0, 0], g['raw'][i]), (name, i)           # ROI samples: bit-exact
        nvalid = np.isfinite(g['raw'][:i + 1][-window:]).sum(axis=0)
        bpm, ptt = res.bpm.cpu().numpy()[0], res.ptt_ms.cpu().numpy()[0]
        joint = int((np.isfinite(g['raw'][:i + 1][-window:]).all(axis=1)).sum())
        py_dev = res.arrays['proc_y'].cpu().numpy()[0]
        rawmax = np.nanmax(np.abs(g['raw'][:i + 1][-window:]), axis=0, initial=1.0)
        resid = [not np.isfinite(py_dev[r]).any() or np.nanmax(np.abs(py_dev[r])) < RESIDUE * rawmax[r] for r in range(2)]
        if any(resid):
            joint = 0
        for r in range(2):
            if nvalid[r] < MIN_N or resid[r]:
                continue       # degenerate warm-up window, see MIN_N
            if ls:
                assert h.same(bpm[r], g['bpm'][i, r]), (name, i, r, bpm[r], g['bpm'][i, r])    # grid frequency: exact
            else:
                assert h.close(bpm[r], g['bpm'][i, r], rtol=1e-12, atol_frac=0), (name, i, r)
        if joint >= MIN_N:
            assert h.close(ptt, g['ptt'][i], rtol=1e-12, atol_frac=0), (name, i, ptt, g['ptt'][i])   # same lag bin
        if i in full_at:
            a = {k: v.cpu().numpy() for k, v in res.arrays.items()}
            # warm-up windows of 2-3 samples filter down to pure rounding residue (~1e-13 of the raw DC level);
            # quantities normalised by that residue's own energy (LS, xcorr) are then noise in the reference too
            residue = resid
            for r in range(2):
                assert h.close(a['proc_x'][0, r], g[f'f{i}_proc_x{r}'], rtol=1e-12, atol_frac=0)
                assert h.close(a['proc_y'][0, r], g[f'f{i}_proc_y{r}'], rtol=1e-4, atol_frac=1e-7, atol=2e-9)
                F = a['num_bins'][0, r]
                assert F == len(g[f'f{i}_spec_x{r}'])
                assert h.close(a['freqs'][0, r, :F], g[f'f{i}_spec_x{r}'], rtol=1e-6, atol_frac=0)
                if not (ls and (nvalid[r] < 4 or residue[r])):
                    assert h.close(a['mags'][0, r, :F], g[f'f{i}_spec_y{r}'], rtol=1e-4, atol_frac=1e-5, atol=1e-30), (name, i, r)
            L = a['num_lags'][0, 0]
            assert L == len(g[f'f{i}_corr_x0'])
            assert h.close(a['lags'][0, 0, :L], g[f'f{i}_corr_x0'], rtol=1e-6, atol_frac=1e-7)
            if not any(residue):
                assert h.close(a['corr'][0, 0, :L], g[f'f{i}_corr_y0'], rtol=1e-4, atol_frac=1e-5)


@pytest.mark.parametrize('windows', ['every_frame', 'last'])
@pytest.mark.parametrize('cfg', [
    dict(channel=orc.CHROM_GREEN, methods=[orc.DETREND_LINEAR, orc.FILTER_FIR], transform=orc.PGRAM_WELCH),   # config 2
    dict(channel=orc.GREEN, methods=[orc.FILTER_BUTTER], transform=orc.PGRAM_LS, min_freq=0.7),              # config 1/5
    dict(channel=orc.GREEN, methods=[orc.INTERP_CUBIC, orc.FILTER_BUTTER], transform=orc.PGRAM_LS),          # config 4
], ids=['c2', 'c1', 'c4'])
def test_multistream_matches_oracle(cfg, windows):
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    cfg = dict(cfg)
    channel, methods, transform = cfg.pop('channel'), cfg.pop('methods'), cfg.pop('transform')
    S, R, W, T, n, H, Wd = 3, 2, 40, 4, 52, 60, 80
    rng = np.random.default_rng(42)
    ts = np.stack([synth.timestamps(rng, n, 30.0, irregular=True, drop=0.05, origin=rng.uniform(0, 100)) for _ in range(S)])
    frames = np.stack([synth.frames(rng, ts[s], H, Wd, f_pulse=1.0 + 0.4 * s) for s in range(S)])
    boxes = np.stack([synth.roi_boxes(rng, n, H, Wd, p_none=0.05, p_oob=0.05) for _ in range(S)])
    eng = BatchedSignalProcessor(S, R, signal_max_samples=W, max_frames_per_step=T, color_channel=channel,
                                 processing_methods=methods, spectrum_transform=transform, windows=windows, **cfg)
    oracles = [orc.OracleStream(R, 1, W, 50, channel, methods, transform, **cfg) for _ in range(S)]
    for g0 in range(0, n, T):
        res = eng.step(torch.from_numpy(frames[:, g0:g0 + T]).cuda(), torch.from_numpy(boxes[:, g0:g0 + T]).cuda(),
                       torch.from_numpy(ts[:, g0:g0 + T].copy()).cuda())
        jobs = res.jobs_per_stream
        bpm = res.bpm.cpu().numpy().reshape(S, jobs, R)
        ptt = res.ptt_ms.cpu().numpy().reshape(S, jobs, 1)
        pidx = res.peak_idx.cpu().numpy().reshape(S, jobs, R)
        lidx = res.lag_idx.cpu().numpy().reshape(S, jobs, 1)
        smp = res.samples.cpu().numpy()
        for s in range(S):
            for j in range(T):
                b = boxes[s, g0 + j]
                rois = [(np.nan,) * 6 if b[r, 0] == synth.NO_BOX else (0, 0, *[int(v) for v in b[r]]) for r in range(R)]
                out = oracles[s].process(frames[s, g0 + j], float(ts[s, g0 + j]), rois)
                assert h.same(smp[s, j], out['samples'])
                jj = j if windows == 'every_frame' else (0 if j == T - 1 else None)
                if jj is None:
                    continue
                nvalid = [int(np.isfinite(oracles[s].raw[r]).sum()) for r in range(R)]
                # windows that filter down to rounding residue (2-3 sample warm-up): derived peaks are noise
                residue = [not np.isfinite(out['proc_y'][r]).any() or
                           np.nanmax(np.abs(out['proc_y'][r])) < RESIDUE * np.nanmax(np.abs(oracles[s].raw[r])) for r in range(R)]
                for r in range(R):
                    if nvalid[r] < MIN_N or residue[r]:
                        continue
                    assert pidx[s, jj, r] == out['peak_idx'][r], (s, g0 + j, r)
                    assert h.close(bpm[s, jj, r], out['bpm'][r], rtol=1e-12, atol_frac=0)
                joint = int((np.isfinite(oracles[s].raw).all(axis=0)).sum())
                if any(residue) or joint < MIN_N:
                    continue
                assert lidx[s, jj, 0] == out['lag_idx'][0], (s, g0 + j)
                assert h.close(ptt[s, jj], out['ptt'], rtol=1e-12, atol_frac=0)


def test_zero_copy_host_frames():
    """F1 reading ROI rows straight from pinned host memory (the e2e path of bench.py)."""
    from bpv import ops, synth
    rng = np.random.default_rng(9)
    N, H, W = 5, 120, 160
    frames = torch.from_numpy(rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)).pin_memory()
    boxes_np = synth.roi_boxes(rng, N, H, W)
    val, sums = ops.roi_sample(frames, torch.from_numpy(boxes_np).cuda(), orc.CHROM_GREEN, want_sums=True)
    torch.cuda.synchronize()
    fn = frames.numpy()
    for f in range(N):
        for r in range(2):
            if boxes_np[f, r, 0] == synth.NO_BOX:
                continue
            assert tuple(int(v) for v in sums[f, r].cpu()) == orc.roi_sums(fn[f], boxes_np[f, r])


@pytest.mark.parametrize('name,cfg', [
    # BASELINE configs[3]: 720p-class 120 fps streams, W = n = 1200, spline resample + band-pass, LS HR, all 2399 xcorr lags
    ('c4_full', dict(W=1200, fps=120.0, methods=[orc.INTERP_CUBIC, orc.FILTER_BUTTER], transform=orc.PGRAM_LS, kw={})),
    # BASELINE configs[2]: irregular timestamps, LS on a 2048-frequency grid, no interpolation
    ('c3_full', dict(W=300, fps=30.0, methods=[], transform=orc.PGRAM_LS, kw=dict(ls_num_freqs=2048))),
    # BASELINE configs[4]/[0]: Butterworth + LS F = n = 300
    ('c5_full', dict(W=300, fps=30.0, methods=[orc.FILTER_BUTTER], transform=orc.PGRAM_LS, kw=dict(min_freq=0.7))),
    ('welch_long', dict(W=1200, fps=120.0, methods=[orc.DETREND_LINEAR, orc.FILTER_FIR], transform=orc.PGRAM_WELCH, kw={})),
    ('dft_long', dict(W=1200, fps=120.0, methods=[orc.FILTER_BUTTER], transform=orc.DFT_RFFT, kw={})),
])
def test_full_size_windows_signals_only(name, cfg):
    """Full BASELINE window sizes through step_signals (steady state: full windows, irregular timestamps, dropped
    detections), every output against the oracle."""
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    S, R, W = 3, 2, cfg['W']
    rng = np.random.default_rng(len(name))
    n = W + 3
    ts = np.stack([synth.timestamps(rng, n, cfg['fps'], irregular=True, drop=0.05) for _ in range(S)])
    ys = np.stack([synth.raw_signals(rng, ts[s], R=R, p_nan=0.02) for s in range(S)])          # [S, R, n]
    eng = BatchedSignalProcessor(S, R, signal_max_samples=W, max_frames_per_step=W, processing_methods=cfg['methods'],
                                 spectrum_transform=cfg['transform'], windows='last', store_arrays=True, **cfg['kw'])
    eng.step_signals(torch.from_numpy(np.ascontiguousarray(ys[:, :, :W].transpose(0, 2, 1))).cuda(), torch.from_numpy(ts[:, :W].copy()).cuda())
    for g in range(W, n):
        res = eng.step_signals(torch.from_numpy(np.ascontiguousarray(ys[:, :, g:g + 1].transpose(0, 2, 1))).cuda(),
                               torch.from_numpy(ts[:, g:g + 1].copy()).cuda())
        a = {k: v.cpu().numpy() for k, v in res.arrays.items()}
        pidx, lidx = res.peak_idx.cpu().numpy(), res.lag_idx.cpu().numpy()
        bpm, ptt = res.bpm.cpu().numpy(), res.ptt_ms.cpu().numpy()
        for s in range(S):
            tw = ts[s, g - W + 1:g + 1]
            proc = [orc.preprocess(tw, ys[s, r, g - W + 1:g + 1], cfg['methods'], **cfg['kw']) for r in range(R)]
            for r in range(R):
                assert h.close(a['proc_y'][s, r], proc[r][1], rtol=1e-6, atol_frac=1e-7), (name, s, r)
                ef, em = orc.spectrum(proc[r][0], proc[r][1], cfg['transform'], **cfg['kw'])
                ex, ey, ei = orc.peak(ef, em)
                F = a['num_bins'][s, r]
                assert F == len(ef)
                assert h.close(a['mags'][s, r, :F], em, rtol=1e-4, atol_frac=1e-5), (name, s, r, np.nanmax(np.abs(a['mags'][s, r, :F] - em)))
                assert pidx[s, r] == ei, (name, s, r, pidx[s, r], ei)
                assert h.close(bpm[s, r], ex * 60, rtol=1e-12, atol_frac=0)
            el, ec = orc.xcorr(proc[0][0], proc[0][1], proc[1][1])
            lx, ly, li = orc.peak(el, ec)
            L = a['num_lags'][s, 0]
            assert L == len(el) and h.close(a['corr'][s, 0, :L], ec, rtol=1e-4, atol_frac=1e-5)
            assert lidx[s, 0] == li and h.close(ptt[s, 0], lx * 1000, rtol=1e-12, atol_frac=0)


def test_three_rois_three_pairs_match_oracle():
    """R = 3 ROI configs (e.g. forehead + cheek + palm, roi.py:24-28) give P = C(3,2) = 3 pairs in
    itertools.combinations order (signal_processor.py:298-299): (0,1), (0,2), (1,2)."""
    from bpv import synth
    from bpv.engine import BatchedSignalProcessor
    S, R, W, T, n, H, Wd = 2, 3, 36, 3, 45, 60, 80
    channel, methods, transform = orc.GREEN, [orc.DETREND_LINEAR, orc.FILTER_BUTTER], orc.PGRAM_LS
    rng = np.random.default_rng(7)
    ts = np.stack([synth.timestamps(rng, n, 30.0, irregular=True, drop=0.05, origin=rng.uniform(0, 100)) for _ in range(S)])
    frames = np.stack([synth.frames(rng, ts[s], H, Wd, f_pulse=1.1 + 0.3 * s) for s in range(S)])
    b2 = np.stack([synth.roi_boxes(rng, n, H, Wd, p_none=0.04, p_oob=0.02) for _ in range(S)])      # [S, n, 2, 4]
    third = b2[:, :, :1].copy()                                  # a cheek-like box: the forehead box shifted down
    ok = third[..., 0] != synth.NO_BOX
    third[..., 1] += np.where(ok, 14, 0); third[..., 3] += np.where(ok, 14, 0)
    boxes = np.ascontiguousarray(np.concatenate([b2[:, :, :1], third, b2[:, :, 1:]], axis=2))          # [S, n, 3, 4]
    eng = BatchedSignalProcessor(S, R, signal_max_samples=W, max_frames_per_step=T, color_channel=channel,
                                 processing_methods=methods, spectrum_transform=transform, windows='every_frame')
    oracles = [orc.OracleStream(R, 1, W, 50, channel, methods, transform) for _ in range(S)]
    assert eng.P == 3
    checked = 0
    for g0 in range(0, n, T):
        res = eng.step(torch.from_numpy(frames[:, g0:g0 + T]).cuda(), torch.from_numpy(boxes[:, g0:g0 + T]).cuda(),
                       torch.from_numpy(ts[:, g0:g0 + T].copy()).cuda())
        pidx = res.peak_idx.cpu().numpy().reshape(S, T, R)
        lidx = res.lag_idx.cpu().numpy().reshape(S, T, 3)
        ptt = res.ptt_ms.cpu().numpy().reshape(S, T, 3)
        rec = res.packed().cpu().numpy().reshape(S, T, 2 * R + 6)
        smp = res.samples.cpu().numpy()
        for s in range(S):
            for j in range(T):
                b = boxes[s, g0 + j]
                rois = [(np.nan,) * 6 if b[r, 0] == synth.NO_BOX else (0, 0, *[int(v) for v in b[r]]) for r in range(R)]
                out = oracles[s].process(frames[s, g0 + j], float(ts[s, g0 + j]), rois)
                assert h.same(smp[s, j], out['samples'])
                raw = oracles[s].raw
                residue = [not np.isfinite(out['proc_y'][r]).any() or
                           np.nanmax(np.abs(out['proc_y'][r])) < RESIDUE * np.nanmax(np.abs(raw[r])) for r in range(R)]
                for r in range(R):
                    if int(np.isfinite(raw[r]).sum()) >= MIN_N and not residue[r]:
                        assert pidx[s, j, r] == out['peak_idx'][r], (s, g0 + j, r)
                for pi, (ra, rb) in enumerate([(0, 1), (0, 2), (1, 2)]):
                    joint = int((np.isfinite(raw[ra]) & np.isfinite(raw[rb])).sum())
                    if joint < MIN_N or residue[ra] or residue[rb]:
                        continue
                    assert lidx[s, j, pi] == out['lag_idx'][pi], (s, g0 + j, pi)
                    assert h.close(ptt[s, j, pi], out['ptt'][pi], rtol=1e-12, atol_frac=0)
                    checked += 1
                # packed record layout: bpm[R] | ptt_ms[P] | peak_idx[R] | lag_idx[P]
                np.testing.assert_array_equal(rec[s, j, R + 3:2 * R + 3], pidx[s, j].astype(np.float64))
                np.testing.assert_array_equal(rec[s, j, 2 * R + 3:], lidx[s, j].astype(np.float64))
    assert checked > 100
