"""GPU parity of the opt-in component output (cvvp_highlight_frames_cc) against what the reference's users compute on
the host in their tracker callback: cv2.connectedComponentsWithStats(bw_frame, connectivity=8) on the highlight mask
(assign_objects_algo.h:124-130 hands the callback bw_frame).  "The same label sets up to canonical relabelling": the
partition must be identical; the GPU numbering is canonical (raster order of the components' first pixels)."""
import cv2
import numpy as np
import pytest

import hl_cases
from oracle import highlight_oracle as ho

pytestmark = pytest.mark.gpu


def _run(ctx, frames, p, max_comps=512):
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)
    try:
        return ctx.highlight_frames_cc(frames, max_comps=max_comps, labels=True)
    finally:
        ctx.highlight_end()


def _check_frame(mask, comps, ncomp, labels):
    n_cv, lab_cv, stats_cv, cent_cv = cv2.connectedComponentsWithStats(mask, connectivity=8, ltype=cv2.CV_32S)
    assert ncomp == n_cv - 1
    assert np.array_equal(labels != 0, mask != 0)
    # identical partition: the (gpu label, cv label) pairs form a bijection
    pairs = np.unique(np.stack([labels.ravel(), lab_cv.ravel()], axis=1), axis=0)
    assert not np.any((pairs[:, 0] == 0) != (pairs[:, 1] == 0))
    pairs = pairs[pairs[:, 0] != 0]
    assert len(pairs) == n_cv - 1
    assert len(np.unique(pairs[:, 0])) == n_cv - 1 and len(np.unique(pairs[:, 1])) == n_cv - 1
    to_cv = dict(map(tuple, pairs))
    prev_first = -1
    for k in range(min(ncomp, comps.shape[0])):
        c = comps[k]
        ys, xs = np.nonzero(labels == k + 1)
        assert c["area"] == len(xs) == stats_cv[to_cv[k + 1], cv2.CC_STAT_AREA]
        assert (c["x0"], c["y0"], c["x1"], c["y1"]) == (xs.min(), ys.min(), xs.max(), ys.max())
        assert c["sum_x"] == xs.sum() and c["sum_y"] == ys.sum()
        cx, cy = cent_cv[to_cv[k + 1]]
        assert abs(c["sum_x"] / c["area"] - cx) < 1e-9 and abs(c["sum_y"] / c["area"] - cy) < 1e-9
        # canonical numbering: components in the raster order of their first pixels
        first = ys[0] * mask.shape[1] + xs[ys == ys[0]].min()
        assert (c["first_y"], c["first_x"]) == divmod(first, mask.shape[1])
        assert first > prev_first
        prev_first = first


@pytest.mark.parametrize("t", range(0, 60, 3))
def test_random_frames(gpu_ctx, t):
    frame, p = hl_cases.random_case(t)
    masks, comps, ncomps, labels = _run(gpu_ctx, frame[None], p)
    assert np.array_equal(masks[0], ho.highlight_objects(frame.copy(), p))
    _check_frame(masks[0], comps[0], int(ncomps[0]), labels[0])


ADV = hl_cases.adversarial_cases()


@pytest.mark.parametrize("case", ADV, ids=[c[0] for c in ADV])
def test_adversarial_frames(gpu_ctx, case):
    _, frame, p = case
    masks, comps, ncomps, labels = _run(gpu_ctx, frame[None], p)
    assert np.array_equal(masks[0], ho.highlight_objects(frame.copy(), p))
    _check_frame(masks[0], comps[0], int(ncomps[0]), labels[0])


def test_batch_of_synthetic_frames_and_stats_only(gpu_ctx):
    """C1-sized synthetic frames in one batch; the label image is optional"""
    from cvvidproc_b200 import synth

    p_ = synth.CONFIG_PARAMS["C1"]
    st = synth.synth_frames(0, 31, p_["width"], p_["height"], p_["seed"], p_["ndisks"])
    bg = np.sort(st, axis=0)[15]
    hp = ho.canonical_params(bg)
    frames = synth.synth_frames(100, 12, p_["width"], p_["height"], p_["seed"], p_["ndisks"])
    masks, comps, ncomps, labels = _run(gpu_ctx, frames, hp, max_comps=64)
    for i in range(frames.shape[0]):
        assert np.array_equal(masks[i], ho.highlight_objects(frames[i].copy(), hp))
        _check_frame(masks[i], comps[i], int(ncomps[i]), labels[i])
    gpu_ctx.highlight_begin(hp.background, np.ascontiguousarray(hp.struct_element), hp.threshold, hp.threshold_lo,
                            hp.threshold_hi, hp.min_size_hyst, hp.min_size_threshold, hp.width_border)
    try:
        m2, c2, n2 = gpu_ctx.highlight_frames_cc(frames, max_comps=64)
    finally:
        gpu_ctx.highlight_end()
    assert np.array_equal(m2, masks) and np.array_equal(n2, ncomps) and np.array_equal(c2, comps)


def test_more_frames_than_the_kernel_keeps_in_flight(gpu_ctx):
    """several chunks through the host-buffer entry point: every scratch slot and staging buffer is reused"""
    rng = np.random.default_rng(8)
    h, w = 48, 96
    bg = np.full((h, w), 180, np.uint8)
    n = 700
    frames = np.repeat(bg[None], n, axis=0)
    for i in range(n):
        for _ in range(1 + i % 4):
            y, x = rng.integers(2, h - 12), rng.integers(2, w - 12)
            frames[i, y : y + rng.integers(4, 10), x : x + rng.integers(4, 10)] = 90
    p = ho.HighlightParams(background=bg, struct_element=np.ones((2, 2), np.uint8), threshold=14, threshold_lo=7,
                           threshold_hi=16, min_size_hyst=4, min_size_threshold=4, width_border=0)
    masks, comps, ncomps, labels = _run(gpu_ctx, frames, p, max_comps=16)
    for i in range(0, n, 23):
        assert np.array_equal(masks[i], ho.highlight_objects(frames[i].copy(), p)), i
        _check_frame(masks[i], comps[i], int(ncomps[i]), labels[i])
    assert ncomps.min() >= 1


def test_more_components_than_slots(gpu_ctx):
    """ncomps reports every component; the first max_comps get statistics, the label image numbers all of them"""
    h, w = 64, 256
    bg = np.full((h, w), 200, np.uint8)
    frame = bg.copy()
    for i in range(20):
        frame[10:20, 10 * i + 3 : 10 * i + 9] = 100  # 20 separate 10x6 blobs
    p = ho.HighlightParams(background=bg, struct_element=np.ones((1, 1), np.uint8), threshold=14, threshold_lo=7,
                           threshold_hi=16, min_size_hyst=0, min_size_threshold=0, width_border=0)
    masks, comps, ncomps, labels = _run(gpu_ctx, frame[None], p, max_comps=5)
    assert np.array_equal(masks[0], ho.highlight_objects(frame.copy(), p))
    assert ncomps[0] == 20 and labels[0].max() == 20
    for k in range(5):
        assert comps[0, k]["area"] == 60 and comps[0, k]["x0"] == 10 * k + 3
