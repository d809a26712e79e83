"""GPU parity of F1 (ROI sampling) through the C-ABI: integer sums and float64 samples must be
bit-exact against the oracle / the reference's golden values."""
import os

import numpy as np
import pytest
import torch

from oracle import bpv_oracle as orc
from tests import helpers as h

pytestmark = pytest.mark.gpu


def _run(frames_np, boxes_np, mode, hint, want_sums=True, offset=0, row_pad=0):
    from bpv import ops
    N, H, W, _ = frames_np.shape
    if offset or row_pad:  # unaligned base address and padded rows (cropped-view case, video_reader.py:101)
        buf = torch.zeros(offset + N * H * (W * 3 + row_pad) + 64, dtype=torch.uint8, device='cuda')
        view = buf[offset:offset + N * H * (W * 3 + row_pad)].view(N, H, W * 3 + row_pad)[:, :, :W * 3].unflatten(2, (W, 3))
        view.copy_(torch.from_numpy(frames_np).cuda())
        frames = view
    else:
        frames = torch.from_numpy(frames_np).cuda()
    boxes = torch.from_numpy(boxes_np.astype(np.int32)).cuda().contiguous()
    val, sums = ops.roi_sample(frames, boxes, mode, want_sums=want_sums, roi_pixels_hint=hint)
    torch.cuda.synchronize()
    return val.cpu().numpy(), None if sums is None else sums.cpu().numpy()


def _check(frames_np, boxes_np, mode, val, sums):
    N, R = boxes_np.shape[:2]
    for f in range(N):
        for r in range(R):
            b = boxes_np[f, r]
            if b[0] == orc.np.iinfo(np.int32).min:
                assert np.isnan(val[f, r])
                continue
            exp = orc.roi_sums(frames_np[f], b)
            if sums is not None:
                assert tuple(int(v) for v in sums[f, r]) == exp, (f, r, b)
            ref = orc.roi_sample(frames_np[f], (0, 0, *[int(v) for v in b]), mode)
            assert h.same(val[f, r], ref), (f, r, b, val[f, r], ref)


@pytest.mark.parametrize('mode', [orc.GREEN, orc.CHROM_GREEN])
def test_golden_values(mode):
    g = np.load(os.path.join(h.GOLDEN, 'roi_sample.npz'))
    rng = np.random.default_rng(int(g['seed']))
    frame = rng.integers(0, 256, (int(g['H']), int(g['W']), 3), dtype=np.uint8)
    boxes = g['boxes'].astype(np.int32)[None]           # one frame, R = len(boxes) ROIs
    for hint in (64, 2000, 100000):
        val, _ = _run(frame[None], boxes, mode, hint)
        assert h.same(val[0], g['values'][mode]), hint


@pytest.mark.parametrize('hint', [100, 4000, 60000])
@pytest.mark.parametrize('shape', [(37, 53), (60, 80), (120, 67)])
def test_random_boxes_exact(hint, shape):
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W + hint)
    N, R = 6, 5
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    boxes = np.stack([rng.integers(-W - 5, W + 9, (N, R)), rng.integers(-H - 5, H + 9, (N, R)),
                      rng.integers(-W - 5, W + 9, (N, R)), rng.integers(-H - 5, H + 9, (N, R))], axis=-1).astype(np.int32)
    boxes[0, 0] = (0, 0, W, H)
    boxes[1, 1] = (np.iinfo(np.int32).min, 0, 0, 0)
    boxes[2, 2] = (3, 3, 3, 9)
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        for offset, pad in ((0, 0), (5, 7), (1, 0)):
            val, sums = _run(frames, boxes, mode, hint, offset=offset, row_pad=pad)
            _check(frames, boxes, mode, val, sums)


def test_full_hd_boxes_exact():
    """BASELINE config-2 geometry: 1080p, forehead/palm-sized boxes incl. out-of-frame ones."""
    from bpv import synth
    H, W, N = 1080, 1920, 4
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    boxes = synth.roi_boxes(rng, N, H, W, p_none=0.2, p_oob=0.3)
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        val, sums = _run(frames, boxes, mode, 96 * 65)
        _check(frames, boxes, mode, val, sums)
    big = np.array([[[0, 0, W, H], [-W, -H, W, H]]] * N, dtype=np.int32)   # max-size ROI: whole frame
    val, sums = _run(frames, big, orc.CHROM_GREEN, H * W)
    _check(frames, big, orc.CHROM_GREEN, val, sums)


def test_unknown_channel_raises():
    from bpv import ops
    frames = torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device='cuda')
    boxes = torch.zeros((1, 1, 4), dtype=torch.int32, device='cuda')
    with pytest.raises(NotImplementedError):
        ops.roi_sample(frames, boxes, 7)


@pytest.mark.parametrize('flip', [False, True])
@pytest.mark.parametrize('shape', [(48, 64), (90, 160)])
def test_videoreader_view_crop_flip_exact(flip, shape):
    """boxes expressed in the reference VideoReader's view (portrait crop frame[:, left:right],
    video_reader.py:97-101, then cv2.flip(frame, 1), :103) sampled straight from the decoded frame must equal the
    reference sampling of the materialised view, bit for bit — including negative / clamped / empty boxes."""
    from bpv import ops
    H, W = shape
    rng = np.random.default_rng(H * 7 + W + int(flip))
    N = 12
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    new_w = int(np.round(H / np.sqrt(2)))                   # video_reader.py:98-100
    left, right = W // 2 - new_w // 2, W // 2 + new_w // 2
    vw = right - left
    boxes = np.empty((N, 2, 4), np.int32)
    boxes[..., 0] = rng.integers(-vw - 5, vw + 5, (N, 2))
    boxes[..., 2] = boxes[..., 0] + rng.integers(-3, vw, (N, 2))
    boxes[..., 1] = rng.integers(-H - 5, H + 5, (N, 2))
    boxes[..., 3] = boxes[..., 1] + rng.integers(-3, H, (N, 2))
    boxes[0, 0] = (0, 0, vw, H)                            # whole view
    boxes[1, 1] = (-10, -10, -2, -2)                       # negative wrap inside the view
    boxes[2, 0] = (orc.np.iinfo(np.int32).min, 0, 0, 0)    # no detection
    src_boxes = ops.view_boxes(torch.from_numpy(boxes).cuda(), vw, H, left, flip)
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        val, sums = ops.roi_sample(torch.from_numpy(frames).cuda(), src_boxes, mode, want_sums=True, roi_pixels_hint=600)
        val, sums = val.cpu().numpy(), sums.cpu().numpy()
        for f in range(N):
            view = frames[f][:, left:right, :]
            if flip:
                view = view[:, ::-1, :]                    # == cv2.flip(view, 1)
            for r in range(2):
                b = boxes[f, r]
                if b[0] == orc.np.iinfo(np.int32).min:
                    assert np.isnan(val[f, r])
                    continue
                ref = orc.roi_sample(np.ascontiguousarray(view), (0, 0, *[int(v) for v in b]), mode)
                assert h.same(val[f, r], ref), (f, r, b, val[f, r], ref)
                sb, sg, sr, n = orc.roi_sums(np.ascontiguousarray(view), b)
                assert tuple(int(v) for v in sums[f, r]) == (sb, sg, sr, n)


@pytest.mark.parametrize('hint', [0, 4000, 400000])
def test_staged_kernel_wide_tall_and_full_frame_boxes(hint):
    """The shared-memory staged F1 kernel (row stride % 16 == 0): ROI rows longer than a CTA's vector columns (> 2 KB),
    ROIs taller than one cp.async batch, 1-pixel-wide columns, the whole frame, unaligned starts."""
    H, W, N = 260, 1920, 3
    rng = np.random.default_rng(77 + hint)
    frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
    boxes = np.array([[(0, 0, W, H), (0, 100, 1920, 110), (5, 0, 6, H), (1, 3, 1919, 259)],
                      [(683, 7, 1367, 255), (1279, 0, 1920, H), (-700, -200, -1, -1), (333, 17, 334, 18)],
                      [(0, 0, 1, 1), (W - 1, H - 1, W, H), (17, 0, 1900, 1), (960, 0, 961, H)]], dtype=np.int32)
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        val, sums = _run(frames, boxes, mode, hint)
        _check(frames, boxes, mode, val, sums)


@pytest.mark.parametrize('shape', [(48, 64), (90, 160), (270, 482)])
def test_nv12_roi_sampling_equals_bgr_of_converted_frame(shape):
    """F1 straight from NV12 planes == the reference sampling of the BGR frame OpenCV would have made
    of them (oracle restatement of cvtColor COLOR_YUV2BGR_NV12, pinned against cv2 in the CPU suite) — integer sums and
    float64 samples bit-exact, Python-slice box semantics, padded pitch."""
    from bpv import ops
    H, W = shape
    rng = np.random.default_rng(H + W)
    N, R, pitch = 5, 4, W + 22
    buf = rng.integers(0, 256, (N, H * 3 // 2, pitch), dtype=np.uint8)
    boxes = np.stack([rng.integers(-W - 3, W + 5, (N, R)), rng.integers(-H - 3, H + 5, (N, R)),
                      rng.integers(-W - 3, W + 5, (N, R)), rng.integers(-H - 3, H + 5, (N, R))], axis=-1).astype(np.int32)
    boxes[0, 0] = (0, 0, W, H)
    boxes[1, 1] = (np.iinfo(np.int32).min, 0, 0, 0)
    boxes[2, 2] = (1, 1, 2, 2)
    boxes[3, 3] = (W - 3, H - 3, W, H)
    bgr = np.stack([orc.nv12_to_bgr(buf[f][:, :W], H, W) for f in range(N)])
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        val, sums = ops.roi_sample_nv12(torch.from_numpy(buf).cuda(), H, W, torch.from_numpy(boxes).cuda(), mode, want_sums=True)
        _check(bgr, boxes, mode, val.cpu().numpy(), sums.cpu().numpy())
        val2, _ = ops.roi_sample_nv12(torch.from_numpy(buf).cuda(), H, W, torch.from_numpy(boxes).cuda(), mode)
        assert h.same(val2.cpu().numpy(), val.cpu().numpy())


@pytest.mark.parametrize('src,dst', [((48, 64), (30, 40)), ((48, 64), (96, 128)), ((90, 160), (45, 80)), ((37, 53), (50, 31)),
                                     ((120, 160), (120, 160))])
def test_resized_roi_sampling_equals_sampling_of_cv2_resize_output(src, dst):
    """F1 with the VideoReader resize (cv2.resize, INTER_LINEAR; video_reader.py:95-96) fused in front —
    boxes in the resized frame, pixels produced on the fly — equals the reference sampling of the resized frame (oracle
    restatement of cv2.resize, pinned against cv2 in the CPU suite): integer sums and float64 samples bit-exact."""
    from bpv import ops
    (sh, sw), (dh, dw) = src, dst
    rng = np.random.default_rng(sh * dw + sw)
    N, R = 4, 4
    frames = rng.integers(0, 256, (N, sh, sw, 3), dtype=np.uint8)
    boxes = np.stack([rng.integers(-dw - 3, dw + 5, (N, R)), rng.integers(-dh - 3, dh + 5, (N, R)),
                      rng.integers(-dw - 3, dw + 5, (N, R)), rng.integers(-dh - 3, dh + 5, (N, R))], axis=-1).astype(np.int32)
    boxes[0, 0] = (0, 0, dw, dh)
    boxes[1, 1] = (np.iinfo(np.int32).min, 0, 0, 0)
    boxes[2, 2] = (dw - 2, dh - 2, dw, dh)
    boxes[3, 3] = (0, 0, 1, 1)
    resized = np.stack([orc.resize_linear_u8(frames[f], dw, dh) for f in range(N)])
    for mode in (orc.GREEN, orc.CHROM_GREEN):
        val, sums = ops.roi_sample_resized(torch.from_numpy(frames).cuda(), dh, dw, torch.from_numpy(boxes).cuda(), mode, want_sums=True)
        _check(resized, boxes, mode, val.cpu().numpy(), sums.cpu().numpy())
