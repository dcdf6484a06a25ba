"""GPU parity of the frame source (csrc/frames.cu behind cvvp_frames_prepare / cvvp_frames_prepare_device /
cvvp_median_push_source) against the cv2 restatement of the reference generator's per-frame work
(oracle/frames_oracle.py; cv_vid_frames_generator_algo.h:141-156).  Bit-exact bytes are the bar."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import frame_cases
from cvvidproc_b200 import _cabi
from oracle import frames_oracle as fo

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((Path(__file__).parent / "golden" / "frames_golden.json").read_text())
CASES = frame_cases.cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_prepare_matches_oracle_and_golden(gpu_ctx, case):
    name, frames, crop, mode = case
    fmt = _cabi.FrameFormat.of(frames.shape[1:], mode, crop)
    got = gpu_ctx.frames_prepare(frames, fmt)
    want = fo.prepare_frames(frames, crop, mode)
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"
    g = next(e for e in GOLDEN if e["name"] == name)
    assert hashlib.sha256(got.tobytes()).hexdigest() == g["output_sha256"]


def test_every_source_alignment_and_width(gpu_ctx):
    """crop_x 0..16 leaves every 16-byte phase of the source segment; widths around the 4-element store granule and
    the 1024-element tile edge"""
    frames = np.random.default_rng(7).integers(0, 256, (2, 6, 1100, 3), dtype=np.uint8)
    for x in range(0, 17):
        for w in (1, 2, 3, 4, 5, 1023, 1024, 1025, 1100 - x):
            if x + w > 1100:
                continue
            for mode in (fo.RGB2GRAY, fo.CHANNEL0):
                fmt = _cabi.FrameFormat.of(frames.shape[1:], mode, (x, 1, w, 4))
                got = gpu_ctx.frames_prepare(frames, fmt)
                assert np.array_equal(got, fo.prepare_frames(frames, (x, 1, w, 4), mode)), (x, w, mode)


def test_strided_source_and_many_frames(gpu_ctx):
    """frame stride larger than a frame (a decoder's padded buffers); more frames than one upload chunk holds"""
    rng = np.random.default_rng(8)
    n, h, w = 70, 600, 800  # 1.44 MB per frame x 70 = 100 MB of decoded bytes > the 64 MB chunk
    buf = rng.integers(0, 256, (n, h * w * 3 + 48), dtype=np.uint8)
    frames = buf[:, : h * w * 3].reshape(n, h, w, 3)
    fmt = _cabi.FrameFormat.of((h, w, 3), fo.RGB2GRAY, (3, 2, 790, 590))
    out = np.empty((n, 590, 790), np.uint8)
    lib = _cabi.load()
    import ctypes

    rc = lib.cvvp_frames_prepare(gpu_ctx.handle, buf.ctypes.data, n, buf.strides[0], ctypes.byref(fmt), out.ctypes.data, 590 * 790)
    assert rc == 0, lib.cvvp_last_error(gpu_ctx.handle)
    c = frames[:, 2:592, 3:793]
    assert np.array_equal(out, fo.rgb2gray_fixed_point(c[..., 0], c[..., 1], c[..., 2]))
    assert np.array_equal(out[5], fo.prepare_frame(frames[5], (3, 2, 790, 590), fo.RGB2GRAY))


def test_device_resident_form_full_hd(gpu_ctx):
    """1080p colour frames resident in HBM -> grey frames in HBM (the form the bench times), whole frame and cropped;
    also from a source pointer that is NOT 16-byte aligned (byte-load path)"""
    import torch

    n, h, w = 8, 1080, 1920
    src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda:0")
    host = src.cpu().numpy()
    for crop in ((0, 0, w, h), (17, 9, 1801, 1000)):
        fmt = _cabi.FrameFormat.of((h, w, 3), fo.RGB2GRAY, crop)
        dst = torch.zeros((n, crop[3] * crop[2]), dtype=torch.uint8, device="cuda:0")
        gpu_ctx.frames_prepare_device(src.data_ptr(), n, h * w * 3, fmt, dst.data_ptr(), dst.stride(0))
        gpu_ctx.synchronize()
        got = dst.cpu().numpy().reshape(n, crop[3], crop[2])
        c = host[:, crop[1]:crop[1] + crop[3], crop[0]:crop[0] + crop[2]]
        assert np.array_equal(got, fo.rgb2gray_fixed_point(c[..., 0], c[..., 1], c[..., 2]))
        assert np.array_equal(got[3], fo.prepare_frame(host[3], crop, fo.RGB2GRAY))
    flat = torch.zeros(n * h * w * 3 + 16, dtype=torch.uint8, device="cuda:0")
    flat[5:5 + n * h * w * 3] = src.reshape(-1)
    fmt = _cabi.FrameFormat.of((h, w, 3), fo.CHANNEL0, (1, 1, 100, 50))
    dst = torch.zeros((n, 5000), dtype=torch.uint8, device="cuda:0")
    gpu_ctx.frames_prepare_device(flat.data_ptr() + 5, n, h * w * 3, fmt, dst.data_ptr(), 5000)
    gpu_ctx.synchronize()
    assert np.array_equal(dst.cpu().numpy().reshape(n, 50, 100), host[:, 1:51, 1:101, 0])


def test_bad_formats_are_loud_errors(gpu_ctx):
    frames = np.zeros((1, 8, 8, 3), np.uint8)
    for fmt in (_cabi.FrameFormat(8, 8, 3, 4, 0, 5, 8, fo.CHANNEL0),   # crop leaves the frame
                _cabi.FrameFormat(8, 8, 3, 0, 0, 8, 8, 7),             # unknown mode
                _cabi.FrameFormat(8, 8, 5, 0, 0, 8, 8, fo.AS_IS),      # 5 channels
                _cabi.FrameFormat(8, 8, 1, 0, 0, 8, 8, fo.RGB2GRAY)):  # grey conversion of one channel
        with pytest.raises(_cabi.CvvpError) as ei:
            gpu_ctx.frames_prepare(frames, fmt)
        assert ei.value.code == -1


@pytest.mark.parametrize("mode", [fo.RGB2GRAY, fo.CHANNEL0, fo.AS_IS])
def test_median_of_decoded_frames_matches_oracle(gpu_ctx, oracle_median, mode):
    """cvvp_median_push_source: decoded colour frames in, median of the prepared frames out == the oracle's median of
    the oracle's prepared frames (GetVideoBackground with crop + grayscale / vid_is_grayscale / colour)"""
    rng = np.random.default_rng(20 + mode)
    n, h, w = 101, 60, 90
    base = rng.integers(0, 256, (h, w, 3), dtype=np.int16)
    frames = np.clip(base[None] + rng.integers(-20, 21, (n, h, w, 3)), 0, 255).astype(np.uint8)
    crop = (5, 7, 77, 40)
    fmt = _cabi.FrameFormat.of((h, w, 3), mode, crop)
    prepared = fo.prepare_frames(frames, crop, mode)
    nelem = int(np.prod(prepared.shape[1:]))
    gpu_ctx.median_begin(nelem, 16)  # small hint: the stack grows while prepared frames are still being written
    for i in range(0, n, 13):
        gpu_ctx.median_push_source(frames[i:i + 13], fmt)
    assert gpu_ctx.median_count() == n
    got = gpu_ctx.median_finish(nelem=nelem).reshape(prepared.shape[1:])
    assert np.array_equal(got, oracle_median(prepared))


def test_colour_median_of_full_width_rows_needs_no_kernel(gpu_ctx, oracle_median):
    """VidBgPack's default on a colour video (no grayscale flag) with a crop that keeps full rows: the copy engine places
    the band in the stack, no preparation kernel is launched, and the median is the oracle's"""
    rng = np.random.default_rng(31)
    n, h, w = 33, 50, 70
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    crop = (0, 6, w, 37)
    fmt = _cabi.FrameFormat.of((h, w, 3), fo.AS_IS, crop)
    prepared = fo.prepare_frames(frames, crop, fo.AS_IS)
    assert np.array_equal(gpu_ctx.frames_prepare(frames, fmt), prepared)
    nelem = 37 * w * 3
    gpu_ctx.median_begin(nelem, n)
    l0 = gpu_ctx.launch_count
    for i in range(0, n, 10):
        gpu_ctx.median_push_source(frames[i:i + 10], fmt)
    assert gpu_ctx.launch_count == l0
    got = gpu_ctx.median_finish(nelem=nelem).reshape(37, w, 3)
    assert np.array_equal(got, oracle_median(prepared))


def test_median_push_source_rejects_a_mismatching_job(gpu_ctx):
    fmt = _cabi.FrameFormat.of((8, 8, 3), fo.RGB2GRAY)
    gpu_ctx.median_begin(65, 4)
    try:
        with pytest.raises(_cabi.CvvpError):
            gpu_ctx.median_push_source(np.zeros((1, 8, 8, 3), np.uint8), fmt)
    finally:
        gpu_ctx.median_abort()
