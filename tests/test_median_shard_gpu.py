"""GPU parity of the frame-sharded median (csrc/median_shard.cu) against the CPU oracle, through the C ABI.

`world` ranks are emulated as `world` contexts of ONE process on ONE device (cvvp_median_shard_attach): every kernel is
the one a real multi-GPU job runs, peer stores simply land in local memory, and the barrier between phases is a
synchronize of every context.  (Real peer memory over NVLink is exercised by `bench.py --gpus N`.)  Bit-exact is the
bar: the result must be sorted[N/2] over ALL ranks' frames (histogram_median_algo.h:160-166).
"""
import os

import numpy as np
import pytest
import torch

from cvvidproc_b200 import _cabi, sharded

pytestmark = pytest.mark.gpu


FORMS = ["two_round", "window", "window_then_full"]
LAST = {}  # what the last _sharded_median call observed (unresolved elements of the one-pass form)


def _sharded_median(frames_per_rank, nelem, form="two_round"):
    """frames_per_rank: list of uint8 arrays (n_r, nelem); returns every rank's result image.
    form: "two_round" = phases 0..3; "window" = the one-pass form (phases 4, 5) and, exactly as ShardedMedian.run does,
    the two-round exchange over the flagged tiles (phases 10..13) only when it left elements undecided;
    "window_then_full" = the same with the whole-image phases 0..3 behind it; "window_only" = phases 4, 5 alone."""
    world = len(frames_per_rank)
    ctxs = [_cabi.Context(0) for _ in range(world)]
    jobs, stacks = [], []
    try:
        stride = (nelem + 127) // 128 * 128
        most = max(fr.shape[0] for fr in frames_per_rank)
        for r, fr in enumerate(frames_per_rank):
            jobs.append(sharded.ShardedMedian(ctxs[r], nelem, r, world, max_rank_frames=most if form != "two_round" else None))
            t = torch.zeros((max(fr.shape[0], 1), stride), dtype=torch.uint8, device="cuda:0")
            if fr.shape[0]:
                t[: fr.shape[0], :nelem] = torch.from_numpy(np.ascontiguousarray(fr)).to("cuda:0")
            stacks.append(t)
        torch.cuda.synchronize()
        sharded.ShardedMedian.connect_local(jobs)

        def walk(phases):
            for p in phases:
                for r, job in enumerate(jobs):
                    job.phase(p, stacks[r].data_ptr(), frames_per_rank[r].shape[0], stride)
                for c in ctxs:
                    c.synchronize()

        LAST.clear()
        if form != "two_round":
            walk((4, 5))
            left = [job.ctx.median_shard_unresolved() for job in jobs]
            assert len(set(left)) == 1, f"ranks disagree on the undecided elements: {left}"
            LAST["unresolved"] = left[0]
        if form == "two_round" or (form == "window_then_full" and LAST["unresolved"] != 0):
            walk(range(4))
        elif form == "window" and LAST["unresolved"] != 0:
            walk(range(10, 14))
        return [job.ctx.copy_to_host(job.result_ptr(), nelem) for job in jobs]
    finally:
        for job in jobs:
            job.close()
        for c in ctxs:
            c.close()


def _split(frames, world, ragged=None):
    n = frames.shape[0]
    parts = []
    for r in range(world):
        first, cnt = sharded.frame_chunk(n, r, world) if ragged is None else ragged[r]
        parts.append(frames[first : first + cnt])
    return parts


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [1, 2, 7, 100, 101, 1000])
def test_random_stack_matches_oracle(oracle_median, world, n, form):
    rng = np.random.default_rng(1000 * world + n)
    h, w = 9, 150
    frames = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    want = oracle_median(frames).reshape(-1)
    outs = _sharded_median(_split(frames.reshape(n, -1), world), h * w, form)
    for r, got in enumerate(outs):
        assert np.array_equal(got, want), f"rank {r} of {world}"
    assert np.array_equal(want, np.sort(frames.reshape(n, -1), axis=0)[n // 2])


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("nelem", [1, 5, 127, 128, 129, 1000, 4099])
def test_ragged_element_counts(oracle_median, nelem, form):
    rng = np.random.default_rng(nelem)
    frames = rng.integers(90, 140, (77, nelem), dtype=np.uint8)
    want = oracle_median(frames.reshape(77, 1, nelem)).reshape(-1)
    for world in (2, 3):
        for got in _sharded_median(_split(frames, world), nelem, form):
            assert np.array_equal(got, want)


@pytest.mark.parametrize("form", FORMS)
def test_uneven_and_empty_chunks(oracle_median, form):
    """ranks may hold different numbers of frames, including none"""
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (50, 700), dtype=np.uint8)
    frames[:, 350:] = rng.integers(120, 126, (50, 350), dtype=np.uint8)  # half of the elements resolve in one pass
    want = oracle_median(frames.reshape(50, 1, 700)).reshape(-1)
    for ragged in ([(0, 50), (50, 0)], [(0, 1), (1, 49)], [(0, 0), (0, 20), (20, 30)], [(0, 33), (33, 0), (33, 17), (50, 0)]):
        for got in _sharded_median(_split(frames, len(ragged), ragged), 700, form):
            assert np.array_equal(got, want)


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("n_local", [513, 1024, 1025, 2049, 4097])
def test_every_tile_variant(oracle_median, n_local, form):
    """local frame counts that select each tile width / buffering mode of the counting kernel (window form: each
    stage count per launch, and several launches = several records per rank)"""
    rng = np.random.default_rng(n_local)
    world = 2
    n = world * n_local - 1
    frames = rng.integers(100, 132, (n, 300), dtype=np.uint8)
    frames[:, 150:] = rng.integers(100, 106, (n, 150), dtype=np.uint8)
    want = oracle_median(frames.reshape(n, 1, 300)).reshape(-1)
    for got in _sharded_median(_split(frames, world), 300, form):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("world,n", [(2, 5001), (3, 7000)])
def test_long_chunks_accumulate_over_launches(oracle_median, world, n):
    """more than 1024 frames per rank: the counting rounds run as several launches whose counts are accumulated
    locally before the last launch pushes them to the owners"""
    rng = np.random.default_rng(n)
    frames = rng.integers(0, 256, (n, 260), dtype=np.uint8)
    frames[:, 0] = 0
    want = oracle_median(frames.reshape(n, 1, 260)).reshape(-1)
    for got in _sharded_median(_split(frames, world), 260):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("stage", ["0", "1"])
@pytest.mark.parametrize("world", [1, 3])
def test_both_push_paths(oracle_median, monkeypatch, stage, world):
    """counting round 1 stores its count vectors directly (32-byte pieces) or staged through shared memory (512 bytes
    per warp); the default picks by world size, CVVP_SHARD_STAGE forces either -- both must give the same image"""
    monkeypatch.setenv("CVVP_SHARD_STAGE", stage)
    rng = np.random.default_rng(17)
    for n, nelem in ((40, 1000), (700, 257), (2300, 130)):  # one stage pair / double buffer / several launches
        frames = rng.integers(0, 256, (n, nelem), dtype=np.uint8)
        frames[:, 0] = 0
        frames[:, nelem // 2:] = rng.integers(30, 36, (n, nelem - nelem // 2), dtype=np.uint8)  # decided by the one pass
        want = oracle_median(frames.reshape(n, 1, nelem)).reshape(-1)
        for form in ("two_round", "window", "window_only"):  # the window records have the same two push paths
            for got in _sharded_median(_split(frames, world), nelem, form):
                if form == "window_only":
                    assert np.array_equal(got[nelem // 2:], want[nelem // 2:]), (stage, world, n, form)
                else:
                    assert np.array_equal(got, want), (stage, world, n, form)


@pytest.mark.parametrize("form", FORMS)
def test_two_valued_split_pins_upper_median(form):
    """exact 50/50 split across ranks: rank 0 holds only the low value, rank 1 only the high one"""
    n = 200
    lo = np.full((n // 2, 256), 16, np.uint8)   # differ in the high nibble
    hi = np.full((n // 2, 256), 32, np.uint8)
    for got in _sharded_median([lo, hi], 256, form):
        assert (got == 32).all()
    for got in _sharded_median([lo, hi[:-1]], 256, form):  # one fewer high value tips it
        assert (got == 16).all()
    lo2 = np.full((n // 2, 256), 0x51, np.uint8)  # same high nibble, differ in the low one
    hi2 = np.full((n // 2, 256), 0x5E, np.uint8)
    for got in _sharded_median([lo2, hi2], 256, form):
        assert (got == 0x5E).all()
    near_lo = np.full((n // 2, 256), 0x51, np.uint8)  # both values inside both windows: decided in one pass
    near_hi = np.full((n // 2, 256), 0x54, np.uint8)
    for got in _sharded_median([near_lo, near_hi], 256, form):
        assert (got == 0x54).all()
    if form.startswith("window"):
        assert LAST["unresolved"] == 0
    for got in _sharded_median([near_lo, near_hi[:-1]], 256, form):
        assert (got == 0x51).all()
    zeros = np.zeros((33, 256), np.uint8)  # value 0 coincides with the zero-filled pad slots
    for got in _sharded_median([zeros, zeros[:5]], 256, form):
        assert (got == 0).all()
    if form.startswith("window"):
        assert LAST["unresolved"] == 0
    ff = np.full((33, 256), 255, np.uint8)
    for got in _sharded_median([ff, zeros[:30]], 256, form):
        assert (got == 255).all()
    for got in _sharded_median([ff, ff[:7]], 256, form):  # window clamped at the top of the range
        assert (got == 255).all()
    if form.startswith("window"):
        assert LAST["unresolved"] == 0


@pytest.mark.parametrize("world", [1, 2, 5])
@pytest.mark.parametrize("n", [640, 3000])
def test_window_form_decides_a_video_background_in_one_pass(oracle_median, world, n):
    """The synthetic stream of SURVEY 8d (smooth background, +-4 noise, moving dark disks): every element is decided
    by phases 4 + 5 alone, and the image is the oracle's.  (Chunks of a few dozen frames are another matter: a disk
    that lingers on a pixel for half of a rank's frames moves that rank's window away -- the elements are reported and
    the two-round exchange settles them, test_short_chunks_of_a_video.)"""
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C2"]
    # two disks on 256x40 pixels cover about the share of the image that C2's thirty cover of a 1080p frame
    frames = synth.synth_frames(0, n, 256, 40, p["seed"], 2).reshape(n, -1)
    want = oracle_median(frames.reshape(n, 1, -1)).reshape(-1)
    for got in _sharded_median(_split(frames, world), frames.shape[1], "window_only"):
        assert np.array_equal(got, want)
    assert LAST["unresolved"] == 0


@pytest.mark.parametrize("world", [2, 5])
def test_short_chunks_of_a_video(oracle_median, world):
    from cvvidproc_b200 import synth

    n = 37
    frames = synth.synth_frames(0, n, 256, 40, 2, 2).reshape(n, -1)
    want = oracle_median(frames.reshape(n, 1, -1)).reshape(-1)
    for got in _sharded_median(_split(frames, world), frames.shape[1], "window"):
        assert np.array_equal(got, want)
    assert LAST["unresolved"] > 0


def test_window_form_reports_what_it_cannot_decide(oracle_median):
    """Chunks with unrelated content: the global median lies outside some rank's window, the element is reported as
    undecided on every rank, and the two-round exchange that follows is exact.  Elements whose chunks agree are decided
    by the one pass even in the same tile."""
    rng = np.random.default_rng(99)
    n, nelem = 300, 640
    frames = rng.integers(100, 104, (n, nelem), dtype=np.uint8)
    frames[: n // 3, :200] = rng.integers(10, 14, (n // 3, 200), dtype=np.uint8)    # rank 0 sees a dark object there
    frames[2 * n // 3 :, 100:300] = rng.integers(200, 230, (n - 2 * n // 3, 200), dtype=np.uint8)
    want = oracle_median(frames.reshape(n, 1, nelem)).reshape(-1)
    parts = _split(frames, 3)
    outs = _sharded_median(parts, nelem, "window_only")
    left = LAST["unresolved"]
    assert 0 < left <= 300
    for got in outs:  # what the one pass did decide is right
        assert np.array_equal(got[300:], want[300:])
    for got in _sharded_median(parts, nelem, "window"):
        assert np.array_equal(got, want)
    assert LAST["unresolved"] == left


@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("n", [2049, 3000, 5001])
def test_long_stack_on_one_gpu_both_forms(oracle_median, gpu_ctx, monkeypatch, window, n):
    """cvvp_median_device beyond 1280 frames: window counting with the device-gated two-pass fallback behind it
    (default) and the two passes alone (CVVP_MEDIAN_WINDOW=0) give the oracle's image; half of the elements are
    uniform noise over the whole range, which the one pass cannot decide, so the gated rounds really run"""
    monkeypatch.setenv("CVVP_MEDIAN_WINDOW", window)
    rng = np.random.default_rng(n)
    nelem = 1000
    frames = rng.integers(60, 66, (n, nelem), dtype=np.uint8)
    want_easy = oracle_median(frames.reshape(n, 1, nelem)).reshape(-1)
    stride = 1024
    t = torch.zeros((n, stride), dtype=torch.uint8, device="cuda:0")
    out = torch.zeros(stride, dtype=torch.uint8, device="cuda:0")
    t[:, :nelem] = torch.from_numpy(frames).to("cuda:0")
    torch.cuda.synchronize()
    gpu_ctx.median_device(t.data_ptr(), n, nelem, stride, out.data_ptr())
    gpu_ctx.synchronize()
    assert np.array_equal(out[:nelem].cpu().numpy(), want_easy)
    frames[:, 500:] = rng.integers(0, 256, (n, 500), dtype=np.uint8)
    want = oracle_median(frames.reshape(n, 1, nelem)).reshape(-1)
    t[:, :nelem] = torch.from_numpy(frames).to("cuda:0")
    torch.cuda.synchronize()
    gpu_ctx.median_device(t.data_ptr(), n, nelem, stride, out.data_ptr())
    gpu_ctx.synchronize()
    assert np.array_equal(out[:nelem].cpu().numpy(), want)


def test_matches_single_gpu_kernel_at_c1_size(gpu_ctx):
    """C1-sized synthetic stack (640x480x100) split over 4 ranks == the single-GPU on-chip select"""
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C1"]
    frames = synth.synth_frames(0, 100, p["width"], 120, p["seed"], p["ndisks"])
    want = gpu_ctx.median(frames).reshape(-1)
    flat = frames.reshape(100, -1)
    for got in _sharded_median(_split(flat, 4), flat.shape[1]):
        assert np.array_equal(got, want)


def test_call_order_errors():
    ctx = _cabi.Context(0)
    try:
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_phase(0)  # no job
        ctx.median_shard_begin(1000, 0, 2)
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_begin(1000, 0, 2)  # already open
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_phase(1)  # peer 1 not mapped
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_import(0, b"\0" * 64)  # own rank
        ctx.median_shard_end()
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_begin(1000, 3, 2)  # rank out of range
        with pytest.raises(_cabi.CvvpError):
            ctx.median_shard_begin(1000, 0, 17)  # too many ranks
    finally:
        ctx.close()


@pytest.mark.parametrize("nelem", [128, 129, 148 * 128 - 1, 148 * 128 + 1, 3 * 148 * 128 + 77])
def test_window_pipeline_stress(oracle_median, nelem):
    """The row-by-row hand-over of the plane buffer (monotonic rows_free counters) and the staged record push under the
    geometries that stress a hand-rolled pipeline: one tile, one tile more / less than there are SMs, every stage count
    that changes the select threads' row count (nst = 1..32 -> 1..8 rows per thread), each repeated -- a race would show
    up as a wrong byte or as the 2 s watchdog trap."""
    rng = np.random.default_rng(nelem)
    for n in (1, 31, 32, 33, 97, 128, 129, 160, 255, 256, 257, 480, 512, 513, 767, 992, 1023, 1024):
        frames = (100 + rng.integers(-3, 5, (n, nelem))).astype(np.uint8)
        want = oracle_median(frames.reshape(n, 1, nelem)).reshape(-1)
        for rep in range(2):
            for world, stage in ((1, "0"), (1, "1"), (2, "1")):
                os.environ["CVVP_SHARD_STAGE"] = stage
                try:
                    outs = _sharded_median(_split(np.concatenate([frames, frames]) if world == 2 else frames, world), nelem, "window_only")
                finally:
                    os.environ.pop("CVVP_SHARD_STAGE", None)
                assert LAST["unresolved"] == 0, (n, world, stage)
                for got in outs:
                    assert np.array_equal(got, want), (n, world, stage, rep)
