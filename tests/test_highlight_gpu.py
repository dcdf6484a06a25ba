"""GPU parity of the highlight stage against the cv2 restatement of highlight_objects_algo.cpp (oracle/highlight_oracle.py),
through the C ABI (cvvp_highlight_begin / cvvp_highlight_frames).  Bit-exact masks are the bar."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import hl_cases
from oracle import highlight_oracle as ho

pytestmark = pytest.mark.gpu


def _gpu(ctx, frames, p):
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)
    try:
        return ctx.highlight_frames(frames)
    finally:
        ctx.highlight_end()


@pytest.mark.parametrize("t", range(60))
def test_random_frames_match_oracle(gpu_ctx, t):
    frame, p = hl_cases.random_case(t)
    got = _gpu(gpu_ctx, frame[None], p)[0]
    want = ho.highlight_objects(frame.copy(), p)
    assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ"


ADV = hl_cases.adversarial_cases()


@pytest.mark.parametrize("case", ADV, ids=[c[0] for c in ADV])
def test_adversarial_frames_match_oracle(gpu_ctx, case):
    _, frame, p = case
    got = _gpu(gpu_ctx, frame[None], p)[0]
    want = ho.highlight_objects(frame.copy(), p)
    assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ"


def test_golden_hashes(gpu_ctx):
    golden = json.loads((Path(__file__).parent / "golden" / "highlight_golden.json").read_text())
    by = {c[0]: c for c in ADV}
    for g in golden:
        if g["kind"] == "adversarial":
            _, frame, p = by[g["name"]]
        elif g["kind"] == "random":
            frame, p = hl_cases.random_case(g["t"])
        else:
            frame, p = hl_cases.synthetic_case(g["cfg"], g["frame_index"], g["scale"])
        got = _gpu(gpu_ctx, frame[None], p)[0]
        assert hashlib.sha256(got.tobytes()).hexdigest() == g["output_sha256"], g["name"]


def test_batch_of_frames_same_as_one_by_one(gpu_ctx):
    """a batch goes through the kernels with blockIdx.z = frame; frames must not influence each other"""
    from cvvidproc_b200 import synth

    w, h, n = 200, 120, 37
    stack = synth.synth_frames(0, n, w, h, 3, 12)
    bg = np.sort(stack[:31], axis=0)[15]
    p = ho.canonical_params(bg)
    got = _gpu(gpu_ctx, stack, p)
    for i in range(n):
        want = ho.highlight_objects(stack[i].copy(), p)
        assert np.array_equal(got[i], want), f"frame {i}"


def test_blobby_frames_larger(gpu_ctx):
    for seed, (h, w) in enumerate([(240, 320), (479, 641), (300, 1000)]):
        frame, bg = hl_cases.blob_frame(h, w, 50 + seed, sigma=4.0, amp=70)
        p = ho.canonical_params(bg)
        got = _gpu(gpu_ctx, frame[None], p)[0]
        want = ho.highlight_objects(frame.copy(), p)
        assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ at {h}x{w}"


def test_full_hd_synthetic_frames(gpu_ctx):
    """BASELINE configs[2] geometry: 1080p frames of the synthetic stream, background = median of the first frames
    computed by the median kernel (stage 2 consumes stage 1's output)."""
    from cvvidproc_b200 import synth

    p_ = synth.CONFIG_PARAMS["C3"]
    w, h = p_["width"], p_["height"]
    stack = synth.synth_frames(0, 33, w, h, p_["seed"], p_["ndisks"])
    bg = gpu_ctx.median(stack)
    assert np.array_equal(bg, np.sort(stack, axis=0)[33 // 2])
    p = ho.canonical_params(bg)
    frames = np.stack([synth.synth_frame(i, w, h, p_["seed"], p_["ndisks"]) for i in (40, 97, 194)])
    got = _gpu(gpu_ctx, frames, p)
    for i in range(frames.shape[0]):
        want = ho.highlight_objects(frames[i].copy(), p)
        assert np.array_equal(got[i], want), f"frame {i}: {(got[i] != want).sum()} pixels differ"


def test_argument_errors(gpu_ctx):
    from cvvidproc_b200 import _cabi

    bg = np.zeros((8, 8), np.uint8)
    with pytest.raises(TypeError):
        gpu_ctx.highlight_begin(bg, np.ones((3, 3), np.int32), 1, 1, 1, 1, 1)
    with pytest.raises(_cabi.CvvpError):  # frames before begin
        lib = _cabi.load()
        rc = lib.cvvp_highlight_frames(gpu_ctx.handle, bg.ctypes.data, 1, 64, bg.ctypes.data, 64)
        if rc != 0:
            raise _cabi.CvvpError(rc, lib.cvvp_last_error(gpu_ctx.handle).decode())
