"""GPU parity of the temporal median against the CPU oracle, through the C ABI (include/cvvp.h).

Bit-exact is the bar (uint8 order statistic).  Mirrors SURVEY.md 9.7's adversarial list for the
median: N in {1,2,3,100,101,255,256,1000}, constant stacks, exact 50/50 two-valued stacks (pins the
UPPER median rule of histogram_median_algo.h:160-166), ragged widths, every tile variant.
"""
import numpy as np
import pytest

from cvvidproc_b200 import _cabi

pytestmark = pytest.mark.gpu


def _rand(rng, n, h, w, lo=0, hi=256):
    return rng.integers(lo, hi, (n, h, w), dtype=np.uint8)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 101, 255, 256, 257, 1000])
def test_random_stack_matches_oracle(gpu_ctx, oracle_median, n):
    rng = np.random.default_rng(n)
    frames = _rand(rng, n, 24, 40)
    got = gpu_ctx.median(frames)
    want = oracle_median(frames)
    assert np.array_equal(got, want)
    assert np.array_equal(got, np.sort(frames, axis=0)[n // 2])


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (3, 5), (479, 641), (16, 128), (17, 129), (2, 4099)])
def test_ragged_geometry(gpu_ctx, oracle_median, shape):
    rng = np.random.default_rng(shape[0] * 10007 + shape[1])
    frames = _rand(rng, 37, *shape)
    assert np.array_equal(gpu_ctx.median(frames), oracle_median(frames))


# every tile variant and both buffering modes.  P=128: N<=512 (two plane buffers) / <=1024 (one buffer); P=64: <=1024
# / <=2048; P=32: <=2048 / <=4096; P=16: <=4096 / <=8192.
@pytest.mark.parametrize("n", [512, 513, 1024, 1025, 2048, 2049, 3000, 4096, 4097, 5000, 8191, 8192])
def test_large_frame_counts(gpu_ctx, oracle_median, n):
    rng = np.random.default_rng(n)
    frames = _rand(rng, n, 5, 77, 90, 140)
    assert np.array_equal(gpu_ctx.median(frames, chunk=512), oracle_median(frames))


@pytest.mark.parametrize("n", [1, 2, 33, 100, 129, 500, 512])
def test_single_buffer_mode_forced(gpu_ctx, oracle_median, n, monkeypatch):
    """CVVP_MEDIAN_BUFFERS=1 routes small frame counts through the one-buffer / 16-select-warp mode as well."""
    monkeypatch.setenv("CVVP_MEDIAN_BUFFERS", "1")
    rng = np.random.default_rng(n + 7)
    frames = _rand(rng, n, 9, 131)
    assert np.array_equal(gpu_ctx.median(frames, chunk=512), oracle_median(frames))


@pytest.mark.parametrize("shape", [(2, 64), (2, 65), (3, 128), (1, 300), (7, 1000)])
def test_tile_pair_tails(gpu_ctx, oracle_median, shape):
    """element counts around tile boundaries (fewer tiles than SMs, partial last tile)"""
    rng = np.random.default_rng(shape[1])
    for n in (5, 600, 1500):
        frames = _rand(rng, n, *shape)
        assert np.array_equal(gpu_ctx.median(frames, chunk=512), oracle_median(frames))


def test_constant_and_two_valued(gpu_ctx, oracle_median):
    const = np.full((100, 8, 16), 200, np.uint8)
    assert np.array_equal(gpu_ctx.median(const), const[0])
    # exact 50/50 split, even N: the upper median is the larger value
    for n in (2, 100, 256, 1000):
        st = np.empty((n, 4, 32), np.uint8)
        st[: n // 2] = 10
        st[n // 2 :] = 250
        got = gpu_ctx.median(st)
        assert (got == 250).all()
        assert np.array_equal(got, oracle_median(st))
        # one fewer high value tips it to the low one
        st[n // 2] = 10
        assert (gpu_ctx.median(st) == 10).all()
    zeros = np.zeros((33, 3, 3), np.uint8)
    assert (gpu_ctx.median(zeros) == 0).all()
    ff = np.full((33, 3, 3), 255, np.uint8)
    assert (gpu_ctx.median(ff) == 255).all()


def test_frame_order_irrelevant_and_multichannel(gpu_ctx, oracle_median):
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (101, 12, 20, 3), dtype=np.uint8)  # 3 interleaved channels, element-wise
    got = gpu_ctx.median(frames)
    assert got.shape == (12, 20, 3)
    assert np.array_equal(got, oracle_median(frames))
    perm = rng.permutation(101)
    assert np.array_equal(gpu_ctx.median(frames[perm]), got)


def test_c1_synthetic_matches_oracle_and_reference(gpu_ctx, oracle_median, ref_median):
    """BASELINE.json configs[0]: 640x480 x 100 synthetic frames, full compare."""
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C1"]
    frames = synth.synth_frames(0, p["nframes"], p["width"], p["height"], p["seed"], p["ndisks"])
    got = gpu_ctx.median(frames)
    assert np.array_equal(got, oracle_median(frames, nthreads=8))
    if ref_median is not None:
        assert np.array_equal(got, ref_median(frames, nthreads=8))


def test_pinned_and_strided_push(gpu_ctx, oracle_median):
    from cvvidproc_b200 import _cabi

    rng = np.random.default_rng(9)
    n, nelem, stride = 50, 1000, 1024
    buf = _cabi.PinnedBuffer(n * stride)
    view = buf.array.reshape(n, stride)
    view[:] = rng.integers(0, 256, (n, stride), dtype=np.uint8)
    gpu_ctx.median_begin(nelem, n)
    gpu_ctx.median_push_raw(view.ctypes.data, n, stride)
    got = gpu_ctx.median_finish(nelem=nelem)
    assert np.array_equal(got, oracle_median(np.ascontiguousarray(view[:, :nelem])))
    assert gpu_ctx.median_last_kernel_ms() > 0
    buf.close()


def test_device_synth_matches_host_synth(gpu_ctx):
    import torch
    from cvvidproc_b200 import synth

    w, h, n, seed, k = 200, 120, 7, 3, 9
    d = torch.empty((n, h * w), dtype=torch.uint8, device="cuda:0")
    gpu_ctx.synth_frames_device(d.data_ptr(), h * w, w, h, 5, n, seed, k)
    gpu_ctx.synchronize()
    torch.cuda.synchronize()
    want = synth.synth_frames(5, n, w, h, seed, k)
    assert np.array_equal(d.cpu().numpy().reshape(n, h, w), want)
    # a row band lands densely
    band = torch.empty((n, 30 * w), dtype=torch.uint8, device="cuda:0")
    gpu_ctx.synth_frames_device(band.data_ptr(), 30 * w, w, h, 5, n, seed, k, row0=40, nrows=30)
    gpu_ctx.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(band.cpu().numpy().reshape(n, 30, w), want[:, 40:70])


def test_device_resident_full_size_property(gpu_ctx):
    """1080p x 1000 (BASELINE configs[1]) at full size: checked through size-independent properties --
    a sampled set of pixels against numpy's exact order statistic, and idempotence under frame reversal."""
    import torch
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C2"]
    w, h, n = p["width"], p["height"], p["nframes"]
    nelem = w * h
    stack = torch.empty((n, nelem), dtype=torch.uint8, device="cuda:0")
    out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
    gpu_ctx.synth_frames_device(stack.data_ptr(), nelem, w, h, 0, n, p["seed"], p["ndisks"])
    gpu_ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
    gpu_ctx.synchronize()
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    cols = np.sort(rng.choice(nelem, 4096, replace=False))
    sample = stack[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    want = np.sort(sample, axis=0)[n // 2]
    got = out.cpu().numpy()
    assert np.array_equal(got[cols], want)
    # torch's own exact k-th value on a contiguous block of columns (independent implementation)
    blk = stack[:, 100000:164000].to(torch.int16)
    kth = torch.kthvalue(blk, n // 2 + 1, dim=0).values.to(torch.uint8)
    assert torch.equal(kth, out[100000:164000])
    # reversing the frame order must not change a single byte
    out2 = torch.empty_like(out)
    rev = torch.flip(stack, dims=[0]).contiguous()
    gpu_ctx.median_device(rev.data_ptr(), n, nelem, nelem, out2.data_ptr())
    gpu_ctx.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("n", [513, 530, 641, 650, 769, 780, 897, 900, 1000, 1024])
def test_every_cta_walks_several_tiles_of_the_single_buffer_mode(gpu_ctx, oracle_median, n):
    """One plane buffer (more than 512 frames), every stage count per select thread (5 .. 8, odd ones split a plane row
    between the slots released first and second), and 469 tiles for 148 CTAs: every CTA goes through even and odd tiles of
    the half-slot pool (csrc/median_pipe.cu), whose slot sets alternate."""
    rng = np.random.default_rng(n * 3 + 1)
    frames = _rand(rng, n, 60, 1000, 60, 200)
    frames[:, 0, :7] = 0                 # value 0 coincides with the zero-filled pad slots
    frames[: n // 2, 1, :5] = 9          # exact split: the upper median
    frames[n // 2 :, 1, :5] = 250
    got = gpu_ctx.median(frames, chunk=256)
    assert np.array_equal(got, oracle_median(frames, nthreads=8))


def test_uhd_geometry_matches_oracle(gpu_ctx, oracle_median):
    """BASELINE configs[4] geometry (3840x2160) at a reduced frame count: full compare against the oracle"""
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C5"]
    frames = synth.synth_frames(0, 41, p["width"], p["height"], p["seed"], p["ndisks"])
    got = gpu_ctx.median(frames, chunk=8)
    assert np.array_equal(got, oracle_median(frames, nthreads=8))


def test_too_many_frames_is_a_loud_error(gpu_ctx):
    """16 x 65535 frames is the limit of the two-pass path (16 slots of 16-bit counts, summed in 32 bits); beyond it
    the call fails, it does not guess"""
    import torch
    from cvvidproc_b200 import _cabi

    n, nelem = 16 * 65535 + 1, 16
    stack = torch.zeros((n, nelem), dtype=torch.uint8, device="cuda:0")
    out = torch.zeros(nelem, dtype=torch.uint8, device="cuda:0")
    with pytest.raises(_cabi.CvvpError) as ei:
        gpu_ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
    assert ei.value.code == -5


@pytest.mark.parametrize("n", [1281, 2049, 3000, 5000, 8193, 10000, 20001, 65535, 65536, 70001, 140000])
def test_long_stacks_take_the_two_pass_path(gpu_ctx, oracle_median, n):
    """more than 1280 frames: counting passes in chunks of <= 1024 frames (csrc/median_shard.cu, one rank)"""
    rng = np.random.default_rng(n)
    nelem = 300 if n < 20000 else 130
    frames = rng.integers(60, 200, (n, 1, nelem), dtype=np.uint8)
    frames[:, 0, 0] = 0          # value 0 coincides with the zero-filled pad slots of every chunk
    frames[: n // 2, 0, 1] = 17  # exact 50/50 split: the UPPER median
    frames[n // 2 :, 0, 1] = 201
    got = gpu_ctx.median(frames, chunk=4096)
    assert np.array_equal(got, oracle_median(frames))
    assert got[0, 1] == 201 and got[0, 0] == 0


@pytest.mark.parametrize("n", [1, 2, 33, 100, 1000, 1024, 1025, 2048])
def test_two_pass_forced_on_short_stacks(gpu_ctx, oracle_median, n, monkeypatch):
    """CVVP_MEDIAN_TWO_PASS=1 routes short stacks through the counting kernels as well (both must agree)"""
    monkeypatch.setenv("CVVP_MEDIAN_TWO_PASS", "1")
    rng = np.random.default_rng(n + 3)
    frames = rng.integers(0, 256, (n, 3, 211), dtype=np.uint8)
    assert np.array_equal(gpu_ctx.median(frames, chunk=512), oracle_median(frames))


@pytest.mark.parametrize("n", [1281, 1500, 2048, 2049, 4096, 4097, 8192])
def test_single_pass_forced_on_long_stacks(gpu_ctx, oracle_median, n, monkeypatch):
    """CVVP_MEDIAN_TWO_PASS=0 keeps the on-chip select with its narrow tile variants covered"""
    monkeypatch.setenv("CVVP_MEDIAN_TWO_PASS", "0")
    rng = np.random.default_rng(n + 5)
    frames = rng.integers(90, 140, (n, 5, 77), dtype=np.uint8)
    assert np.array_equal(gpu_ctx.median(frames, chunk=512), oracle_median(frames))


def test_cuda_reproduces_reference_golden(gpu_ctx):
    """tests/golden/median_golden.json holds outputs of the reference's own class; the CUDA path must hit every
    hash except the two cases that force the reference's u8 counters to saturate (unreachable through
    GetVideoBackground, SURVEY.md 8a a5 -- the CUDA path is an exact rank select)."""
    import hashlib
    import importlib.util
    import json
    from pathlib import Path

    gdir = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("make_golden", gdir / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for case in json.loads((gdir / "median_golden.json").read_text()):
        if case.get("bin_bytes"):
            continue
        frames = mg.case_input(case)
        got = gpu_ctx.median(frames)
        assert hashlib.sha256(got.tobytes()).hexdigest() == case["output_sha256"], case["name"]


# ---- constant-memory form: a stack that cannot grow folds into value histograms (csrc/median_hist.cu) -------------------
@pytest.mark.parametrize("n,cap,chunk", [(1000, 256, 100), (300, 16, 7), (65, 64, 64), (64, 64, 10), (2100, 1000, 512),
                                         (5000, 333, 1024), (40, 1, 3), (200, 3, 50)])
def test_stack_that_cannot_grow_folds_into_value_histograms(gpu_ctx, oracle_median, monkeypatch, n, cap, chunk):
    """CVVP_MEDIAN_RESIDENT_MAX caps the resident stack the way a full device does: the job folds the resident frames
    into per-element value histograms and goes on -- memory independent of the frame count, like the reference's
    HistogramMedianAlgo (histogram_median_algo.h:116-193) -- and the result stays the upper median of ALL frames"""
    monkeypatch.setenv("CVVP_MEDIAN_RESIDENT_MAX", str(cap))
    rng = np.random.default_rng(n * 7 + cap)
    frames = rng.integers(80, 150, (n, 9, 131), dtype=np.uint8)
    frames[:, 0, :5] = 0                       # constant columns at both ends of the value range
    frames[:, 0, 5:9] = 255
    frames[: n // 2, 1, :6] = 9                # exact split: the upper median
    frames[n // 2 :, 1, :6] = 250
    frames[:, 2, :] = rng.integers(0, 256, (n, 131), dtype=np.uint8)   # the whole value range
    nelem = 9 * 131
    gpu_ctx.median_begin(nelem, n)
    for i in range(0, n, chunk):
        gpu_ctx.median_push(frames[i:i + chunk])
        assert gpu_ctx.median_count() == min(i + chunk, n)
    if n > cap:
        with pytest.raises(_cabi.CvvpError) as ei:   # the stack no longer holds every frame
            gpu_ctx.median_stack_device()
        assert ei.value.code == -4 and "folded" in str(ei.value)
    got = gpu_ctx.median_finish(nelem=nelem).reshape(9, 131)
    want = oracle_median(frames)
    assert np.array_equal(got, want)
    assert (got[1, :6] == 250).all()
    # the context goes back to resident jobs afterwards
    monkeypatch.delenv("CVVP_MEDIAN_RESIDENT_MAX")
    assert np.array_equal(gpu_ctx.median(frames[:100]), oracle_median(frames[:100]))


def test_folding_with_device_prepared_colour_frames_and_pinned_sources(gpu_ctx, oracle_median, monkeypatch):
    """the same through cvvp_median_push_source (decoded colour frames, cropped and converted on the device) and from
    page-locked host memory"""
    from oracle import frames_oracle as fo

    monkeypatch.setenv("CVVP_MEDIAN_RESIDENT_MAX", "40")
    rng = np.random.default_rng(77)
    n, h, w = 203, 50, 70
    base = rng.integers(0, 256, (h, w, 3), dtype=np.int16)
    frames = np.clip(base[None] + rng.integers(-25, 26, (n, h, w, 3)), 0, 255).astype(np.uint8)
    for mode, crop in ((fo.RGB2GRAY, (3, 4, 61, 40)), (fo.AS_IS, (3, 4, 61, 40)), (fo.AS_IS, (0, 6, 70, 31))):  # the last one:
        fmt = _cabi.FrameFormat.of((h, w, 3), mode, crop)                          # full-width rows, copied without a kernel
        prepared = fo.prepare_frames(frames, crop, mode)
        nelem = int(np.prod(prepared.shape[1:]))
        gpu_ctx.median_begin(nelem, n)
        for i in range(0, n, 33):
            gpu_ctx.median_push_source(frames[i:i + 33], fmt)
        assert gpu_ctx.median_count() == n
        got = gpu_ctx.median_finish(nelem=nelem).reshape(prepared.shape[1:])
        assert np.array_equal(got, oracle_median(prepared)), mode
    pinned = _cabi.PinnedBuffer(n * h * w * 3)
    host = pinned.array[: n * h * w * 3].reshape(n, h, w, 3)
    host[:] = frames
    gpu_ctx.median_begin(h * w * 3, n)
    gpu_ctx.median_push(host[:150])
    gpu_ctx.median_push(host[150:])
    got = gpu_ctx.median_finish(nelem=h * w * 3).reshape(h, w, 3)
    assert np.array_equal(got, oracle_median(frames))


def test_a_hint_beyond_the_device_leaves_a_smaller_stack(oracle_median):
    """GetVideoBackground announces frames_to_analyze; a video longer than the device is wide (200 000 frames of 1080p:
    415 GB) must not fail at cvvp_median_begin: the job takes a stack that fits and folds as it goes"""
    from cvvidproc_b200 import synth

    p = synth.CONFIG_PARAMS["C2"]
    frames = synth.synth_frames(0, 40, p["width"], p["height"], p["seed"], p["ndisks"])
    with _cabi.Context(0) as ctx:
        ctx.median_begin(p["width"] * p["height"], 200_000)
        ctx.median_push(frames[:25])
        ctx.median_push(frames[25:])
        assert ctx.median_count() == 40
        got = ctx.median_finish(nelem=p["width"] * p["height"]).reshape(p["height"], p["width"])
    assert np.array_equal(got, oracle_median(frames, nthreads=8))
