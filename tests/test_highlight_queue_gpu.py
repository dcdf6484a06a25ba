"""GPU tests of the asynchronous ordered highlight queue (cvvp_highlight_queue_begin / submit / ready / next / end):
the reference's bounded token queues and in-order hand-over (token_queue.h:209-214, mat_set_intermediary.h:50-68)
rebuilt on streams and events.  Results must be the synchronous path's and the cv2 oracle's, in submission order."""
import numpy as np
import pytest

import hl_cases
from cvvidproc_b200 import _cabi, synth
from oracle import frames_oracle as fo
from oracle import highlight_oracle as ho

pytestmark = pytest.mark.gpu


def _stream(cfg="C1", first=100, n=23, height=96):
    p = synth.CONFIG_PARAMS[cfg]
    frames = synth.synth_frames(first, n, p["width"], height, p["seed"], p["ndisks"])
    bg = np.median(synth.synth_frames(0, 31, p["width"], height, p["seed"], p["ndisks"]), axis=0).astype(np.uint8)
    return frames, ho.canonical_params(bg)


def _begin(ctx, p):
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)


@pytest.mark.parametrize("depth,batch", [(1, 4), (2, 5), (3, 7), (4, 1)])
def test_queue_returns_batches_in_order_and_matches_oracle(gpu_ctx, depth, batch):
    frames, p = _stream()
    want = np.stack([ho.highlight_objects(f.copy(), p) for f in frames])
    _begin(gpu_ctx, p)
    try:
        sync = gpu_ctx.highlight_frames(frames)
        gpu_ctx.highlight_queue_begin(depth, batch)
        got, sizes, i = [], [], 0
        while i < len(frames) or gpu_ctx.highlight_queue_pending():
            while i < len(frames) and gpu_ctx.highlight_queue_pending() < depth:
                chunk = frames[i:i + batch].copy()
                gpu_ctx.highlight_submit(chunk)
                chunk[:] = 0  # the caller's buffer is free as soon as submit returns (a moved token)
                sizes.append(min(batch, len(frames) - i))
                i += batch
            m = gpu_ctx.highlight_next()
            assert m.shape[0] == sizes[len(got)]
            got.append(m.copy())
        gpu_ctx.highlight_queue_end()
    finally:
        gpu_ctx.highlight_end()
    got = np.concatenate(got)
    assert np.array_equal(got, sync)
    assert np.array_equal(got, want)


def test_back_pressure_and_state_errors(gpu_ctx):
    frames, p = _stream(n=6)
    _begin(gpu_ctx, p)
    try:
        with pytest.raises(_cabi.CvvpError):  # not begun
            gpu_ctx.highlight_submit(frames[:1])
        gpu_ctx.highlight_queue_begin(2, 2)
        with pytest.raises(_cabi.CvvpError):  # begun twice
            gpu_ctx.highlight_queue_begin(2, 2)
        with pytest.raises(_cabi.CvvpError):  # nothing pending
            gpu_ctx.highlight_next()
        assert not gpu_ctx.highlight_queue_ready()
        with pytest.raises(_cabi.CvvpError):  # more frames than a slot holds
            gpu_ctx.highlight_submit(frames[:3])
        gpu_ctx.highlight_submit(frames[0:2])
        gpu_ctx.highlight_submit(frames[2:4])
        assert gpu_ctx.highlight_queue_pending() == 2
        with pytest.raises(_cabi.CvvpError) as ei:  # token_storage_limit reached
            gpu_ctx.highlight_submit(frames[4:6])
        assert ei.value.code == -4
        a = gpu_ctx.highlight_next().copy()
        gpu_ctx.highlight_submit(frames[4:6])
        b = gpu_ctx.highlight_next().copy()
        gpu_ctx.synchronize()
        assert gpu_ctx.highlight_queue_ready()
        c = gpu_ctx.highlight_next().copy()
        assert gpu_ctx.highlight_queue_pending() == 0
        want = np.stack([ho.highlight_objects(f.copy(), p) for f in frames])
        assert np.array_equal(np.concatenate([a, b, c]), want)
        # pending results may be dropped
        gpu_ctx.highlight_submit(frames[0:2])
        gpu_ctx.highlight_queue_end()
        assert gpu_ctx.highlight_queue_pending() == 0
    finally:
        gpu_ctx.highlight_end()


@pytest.mark.parametrize("mode", [fo.RGB2GRAY, fo.CHANNEL0])
def test_queue_with_decoded_colour_frames(gpu_ctx, mode):
    """decoded 3-channel frames with a crop go through the device frame preparation inside the queue: the masks are
    the oracle's masks of the oracle's prepared frames"""
    rng = np.random.default_rng(5 + mode)
    h, w, n = 70, 120, 11
    crop = (9, 4, 100, 60)
    grey_bg = np.full((h, w), 150, np.uint8)
    frames = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        f, _ = hl_cases.blob_frame(h, w, 900 + i, sigma=2.0, amp=70)
        frames[i] = np.stack([f, np.roll(f, 3, axis=1), rng.integers(0, 256, (h, w), dtype=np.uint8)], axis=-1)
    frames[..., 0] = np.minimum(frames[..., 0], grey_bg)
    prepared = fo.prepare_frames(frames, crop, mode)
    bg = np.median(prepared, axis=0).astype(np.uint8)
    p = ho.canonical_params(bg)
    want = np.stack([ho.highlight_objects(f.copy(), p) for f in prepared])
    _begin(gpu_ctx, p)
    try:
        fmt = _cabi.FrameFormat.of((h, w, 3), mode, crop)
        gpu_ctx.highlight_queue_begin(2, 4, fmt)
        got = []
        for i in range(0, n, 4):
            if gpu_ctx.highlight_queue_pending() == 2:
                got.append(gpu_ctx.highlight_next().copy())
            gpu_ctx.highlight_submit(frames[i:i + 4])
        while gpu_ctx.highlight_queue_pending():
            got.append(gpu_ctx.highlight_next().copy())
        with pytest.raises(_cabi.CvvpError):  # a format whose output is not the background's geometry
            gpu_ctx.highlight_queue_end()
            gpu_ctx.highlight_queue_begin(2, 4, _cabi.FrameFormat.of((h, w, 3), mode, (0, 0, 50, 60)))
    finally:
        gpu_ctx.highlight_end()
    assert np.array_equal(np.concatenate(got), want)


def test_queue_with_components(gpu_ctx):
    frames, p = _stream(n=9)
    _begin(gpu_ctx, p)
    try:
        m_sync, c_sync, n_sync = gpu_ctx.highlight_frames_cc(frames, max_comps=64)
        gpu_ctx.highlight_queue_begin(3, 4, None, 64)
        for i in range(0, 9, 4):
            gpu_ctx.highlight_submit(frames[i:i + 4])
        ms, cs, ns = [], [], []
        while gpu_ctx.highlight_queue_pending():
            m, c, k = gpu_ctx.highlight_next()
            ms.append(m.copy()), cs.append(c.copy()), ns.append(k.copy())
    finally:
        gpu_ctx.highlight_end()
    assert np.array_equal(np.concatenate(ms), m_sync)
    assert np.array_equal(np.concatenate(ns), n_sync)
    cs = np.concatenate(cs)
    for f in range(9):
        assert np.array_equal(cs[f, :n_sync[f]], c_sync[f, :n_sync[f]])


def test_full_hd_queue_matches_synchronous_path(gpu_ctx):
    p3 = synth.CONFIG_PARAMS["C3"]
    frames = synth.synth_frames(300, 40, p3["width"], p3["height"], p3["seed"], p3["ndisks"])
    bg = np.median(synth.synth_frames(0, 15, p3["width"], p3["height"], p3["seed"], p3["ndisks"]), axis=0).astype(np.uint8)
    p = ho.canonical_params(bg)
    _begin(gpu_ctx, p)
    try:
        sync = gpu_ctx.highlight_frames(frames)
        gpu_ctx.highlight_queue_begin(3, 16)
        got = []
        for i in range(0, 40, 16):
            gpu_ctx.highlight_submit(frames[i:i + 16])
        while gpu_ctx.highlight_queue_pending():
            got.append(gpu_ctx.highlight_next().copy())
    finally:
        gpu_ctx.highlight_end()
    got = np.concatenate(got)
    assert np.array_equal(got, sync)
    assert np.array_equal(got[7], ho.highlight_objects(frames[7].copy(), p))


@pytest.mark.parametrize("depth,batch,decoded", [(1, 3, False), (3, 5, False), (2, 4, True)])
def test_zero_copy_slots_match_the_copying_forms(gpu_ctx, depth, batch, decoded):
    """cvvp_highlight_slot_acquire / slot_commit / next_view / view_release: the producer writes whole frames straight
    into the slot's pinned input (decoded colour frames with a crop when the queue has a format) and the consumer reads
    the slot's pinned results; same masks, same order as submit / next; the two forms can be mixed"""
    frames, p = _stream(n=17)
    H, W = frames.shape[1:]
    fmt = None
    src = frames
    if decoded:
        pad = np.zeros((len(frames), H + 6, W + 10, 3), np.uint8)
        pad[:, 4:4 + H, 7:7 + W, 0] = frames
        pad[..., 1] = 255 - pad[..., 0]
        src = pad
        fmt = _cabi.FrameFormat.of(pad.shape[1:], _cabi.FRAMES_CHANNEL0, (7, 4, W, H))
    want = np.stack([ho.highlight_objects(f.copy(), p) for f in frames])
    _begin(gpu_ctx, p)
    try:
        gpu_ctx.highlight_queue_begin(depth, batch, fmt)
        fb = int(np.prod(src.shape[1:]))
        got, i, k = [], 0, 0
        while i < len(frames) or gpu_ctx.highlight_queue_pending():
            while i < len(frames) and gpu_ctx.highlight_queue_pending() < depth:
                n = min(batch, len(frames) - i)
                if k % 3 == 2:  # every third batch through the copying form
                    gpu_ctx.highlight_submit(src[i:i + n])
                else:
                    slot = gpu_ctx.highlight_slot_acquire()
                    assert slot.shape[0] == batch and slot.shape[1] >= fb
                    slot[:n, :fb] = src[i:i + n].reshape(n, fb)
                    gpu_ctx.highlight_slot_commit(n)
                i += n
                k += 1
            if len(got) % 2:
                got.append(gpu_ctx.highlight_next().copy())
            else:
                view = gpu_ctx.highlight_next_view()
                with pytest.raises(_cabi.CvvpError):  # the oldest batch is lent out
                    gpu_ctx.highlight_next()
                got.append(np.array(view))
                gpu_ctx.highlight_view_release()
        # a decoder working ahead: every free slot may be out at once, they are committed in acquisition order, and the
        # end of the stream hands the rest back unused
        slots = [gpu_ctx.highlight_slot_acquire() for _ in range(depth)]
        with pytest.raises(_cabi.CvvpError):  # the ring is full
            gpu_ctx.highlight_slot_acquire()
        m = min(batch, 3)
        for j, sl in enumerate(slots):
            sl[:m, :fb] = src[j:j + m].reshape(m, fb)
        gpu_ctx.highlight_slot_commit(m)
        for _ in slots[1:]:
            gpu_ctx.highlight_slot_commit(0)
        assert gpu_ctx.highlight_queue_pending() == 1
        assert np.array_equal(gpu_ctx.highlight_next(), want[:m])
        if depth >= 2:
            a, b = gpu_ctx.highlight_slot_acquire(), gpu_ctx.highlight_slot_acquire()
            gpu_ctx.highlight_slot_commit(0)
            with pytest.raises(_cabi.CvvpError):  # the slot after a returned one cannot be queued any more
                gpu_ctx.highlight_slot_commit(1)
            gpu_ctx.highlight_slot_commit(0)
        assert gpu_ctx.highlight_queue_pending() == 0
        with pytest.raises(_cabi.CvvpError):
            gpu_ctx.highlight_view_release()
        gpu_ctx.highlight_queue_end()
    finally:
        gpu_ctx.highlight_end()
    assert np.array_equal(np.concatenate(got), want)


def test_buffers_of_a_finished_job_are_reused_and_trimmed(gpu_ctx):
    """the queue's pinned slots / device buffers and the fused kernel's scratch are parked when a job ends and handed to
    the next job of the same geometry (csrc/pool.hpp); results are unaffected and cvvp_pool_trim releases what is parked"""
    lib = _cabi.load()
    frames, p = _stream(n=10)
    want = np.stack([ho.highlight_objects(f.copy(), p) for f in frames])
    lib.cvvp_pool_trim()
    for _ in range(3):  # the second and third job run on recycled (dirty) buffers
        _begin(gpu_ctx, p)
        try:
            gpu_ctx.highlight_queue_begin(2, 64)  # 64 frames x 61 KB: above the pool's 1 MB threshold
            gpu_ctx.highlight_submit(frames[:7])
            gpu_ctx.highlight_submit(frames[7:])
            got = np.concatenate([gpu_ctx.highlight_next().copy(), gpu_ctx.highlight_next().copy()])
            gpu_ctx.highlight_queue_end()
        finally:
            gpu_ctx.highlight_end()
        assert np.array_equal(got, want)
    assert lib.cvvp_pool_trim() > 0
    assert lib.cvvp_pool_trim() == 0
