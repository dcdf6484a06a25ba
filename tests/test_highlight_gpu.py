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


FUSED, PIXEL = 0, 1  # cvvp_highlight_set_path: the fused persistent-CTA kernel (default) / the per-pixel kernels


def _gpu(ctx, frames, p, path=FUSED):
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)
    try:
        ctx.highlight_set_path(path)
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


@pytest.mark.parametrize("t", range(0, 60, 7))
def test_pixel_path_random_frames_match_oracle(gpu_ctx, t):
    """the per-pixel kernels stay held to the oracle: they are the on-device cross-check of the fused kernel"""
    frame, p = hl_cases.random_case(t)
    got = _gpu(gpu_ctx, frame[None], p, PIXEL)[0]
    assert np.array_equal(got, ho.highlight_objects(frame.copy(), p))


@pytest.mark.parametrize("case", ADV[::3], ids=[c[0] for c in ADV[::3]])
def test_pixel_path_adversarial_frames_match_oracle(gpu_ctx, case):
    _, frame, p = case
    got = _gpu(gpu_ctx, frame[None], p, PIXEL)[0]
    assert np.array_equal(got, ho.highlight_objects(frame.copy(), p))


@pytest.mark.parametrize("hw", [(40, 32), (33, 64), (50, 96), (21, 128), (30, 160), (24, 288), (17, 640), (9, 1056)])
def test_word_aligned_widths_match_oracle(gpu_ctx, hw):
    """W % 32 == 0 takes the 128-bit load/store path of the fused kernel; row pitches that are / are not powers of two"""
    h, w = hw
    for seed in range(3):
        frame, bg = hl_cases.blob_frame(h, w, 300 + seed, sigma=2.0, amp=70)
        for p in (ho.canonical_params(bg),
                  ho.HighlightParams(bg, np.ones((1, 1), np.uint8), 12, 6, 20, 0, 3, 5),
                  ho.HighlightParams(bg, np.ones((3, 2), np.uint8), -1, 9, 14, 10, 30, 5)):
            got = _gpu(gpu_ctx, frame[None], p)[0]
            want = ho.highlight_objects(frame.copy(), p)
            assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ at {h}x{w} seed {seed}"


@pytest.mark.parametrize("seed", range(6))
def test_noise_frames_match_oracle(gpu_ctx, seed):
    """salt-and-pepper difference images: the run count approaches the pixel count (worst case of the run-based
    labelling), nesting is deep and almost every contour is small"""
    rng = np.random.default_rng(7000 + seed)
    h, w = int(rng.integers(20, 80)), int(rng.choice([31, 64, 75, 96]))
    bg = np.full((h, w), 128, np.uint8)
    frame = (128 - rng.integers(0, 40, (h, w))).astype(np.uint8)
    one = np.ones((1, 1), np.uint8)
    for p in (ho.HighlightParams(bg, one, 20, 10, 30, int(rng.choice([0, 2, 6])), int(rng.choice([0, 2, 6])), 5),
              ho.HighlightParams(bg, one, 5, 30, 8, 4, 0, 5),
              ho.canonical_params(bg)):
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


def test_fused_and_pixel_paths_agree_at_full_hd(gpu_ctx):
    """device vs device at the BASELINE geometry, more frames than the kernel keeps in flight (the frame queue wraps and
    every scratch slot is reused), plus two frames of pure noise"""
    from cvvidproc_b200 import synth

    p_ = synth.CONFIG_PARAMS["C3"]
    w, h = p_["width"], p_["height"]
    stack = synth.synth_frames(0, 15, w, h, p_["seed"], p_["ndisks"])
    bg = np.sort(stack, axis=0)[7]
    p = ho.canonical_params(bg)
    gpu_ctx.highlight_begin(p.background, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                            p.min_size_threshold, p.width_border)
    in_flight = gpu_ctx.highlight_frames_in_flight()
    gpu_ctx.highlight_end()
    assert in_flight >= 1
    n = min(in_flight + 37, 420)
    frames = np.stack([synth.synth_frame(1000 + 3 * i, w, h, p_["seed"], p_["ndisks"]) for i in range(n)])
    rng = np.random.default_rng(5)
    frames[5] = (bg.astype(np.int32) - rng.integers(0, 30, bg.shape)).clip(0, 255).astype(np.uint8)
    frames[n - 2] = (bg.astype(np.int32) - rng.integers(0, 18, bg.shape)).clip(0, 255).astype(np.uint8)
    fused = _gpu(gpu_ctx, frames, p, FUSED)
    pixel = _gpu(gpu_ctx, frames, p, PIXEL)
    for i in range(n):
        assert np.array_equal(fused[i], pixel[i]), f"frame {i}: {(fused[i] != pixel[i]).sum()} pixels differ"
    for i in (0, 5, n - 1):
        assert np.array_equal(fused[i], ho.highlight_objects(frames[i].copy(), p)), f"frame {i} vs oracle"


def test_bubble_video_geometry_many_frames(gpu_ctx):
    """BASELINE configs[3] geometry: 512x256 frames, several times more frames than the kernel keeps in flight; every
    frame fused vs per-pixel path, every 17th vs the oracle"""
    from cvvidproc_b200 import synth

    p_ = synth.CONFIG_PARAMS["C4"]
    w, h = p_["width"], p_["height"]
    stack = synth.synth_frames(0, 101, w, h, p_["seed"], p_["ndisks"])
    bg = gpu_ctx.median(stack)
    p = ho.canonical_params(bg)
    frames = synth.synth_frames(500, 700, w, h, p_["seed"], p_["ndisks"])
    fused = _gpu(gpu_ctx, frames, p, FUSED)
    pixel = _gpu(gpu_ctx, frames, p, PIXEL)
    assert np.array_equal(fused, pixel)
    for i in range(0, 700, 17):
        assert np.array_equal(fused[i], ho.highlight_objects(frames[i].copy(), p)), f"frame {i} vs oracle"


# The fused kernel is built twice (csrc/highlight_fused.cu): 1024-thread CTAs, one per SM ("large", frames above 512x512)
# and 256-thread CTAs, four per SM ("small").  The cases above run on whichever the geometry selects -- the small frames
# on "small", the 1080p ones on "large"; the tests below force the other build through CVVP_HL_VARIANT.
@pytest.mark.parametrize("t", range(0, 60, 2))
def test_large_variant_random_frames_match_oracle(gpu_ctx, t, monkeypatch):
    monkeypatch.setenv("CVVP_HL_VARIANT", "large")
    frame, p = hl_cases.random_case(t)
    got = _gpu(gpu_ctx, frame[None], p)[0]
    want = ho.highlight_objects(frame.copy(), p)
    assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ"


@pytest.mark.parametrize("case", ADV, ids=[c[0] for c in ADV])
def test_large_variant_adversarial_frames_match_oracle(gpu_ctx, case, monkeypatch):
    monkeypatch.setenv("CVVP_HL_VARIANT", "large")
    _, frame, p = case
    got = _gpu(gpu_ctx, frame[None], p)[0]
    assert np.array_equal(got, ho.highlight_objects(frame.copy(), p))


def test_variants_agree_on_both_baseline_geometries(gpu_ctx, monkeypatch):
    """512x256 (BASELINE configs[3]) forced onto the 1024-thread build and 1080p (configs[2]) forced onto the 256-thread
    build: same masks as the build the geometry selects, the noise frames included"""
    from cvvidproc_b200 import synth

    for cfg, n, other in (("C4", 700, "large"), ("C3", 40, "small")):
        p_ = synth.CONFIG_PARAMS[cfg]
        w, h = p_["width"], p_["height"]
        bg = np.sort(synth.synth_frames(0, 15, w, h, p_["seed"], p_["ndisks"]), axis=0)[7]
        p = ho.canonical_params(bg)
        frames = synth.synth_frames(2000, n, w, h, p_["seed"], p_["ndisks"])
        rng = np.random.default_rng(9)
        frames[3] = (bg.astype(np.int32) - rng.integers(0, 30, bg.shape)).clip(0, 255).astype(np.uint8)
        monkeypatch.delenv("CVVP_HL_VARIANT", raising=False)
        auto = _gpu(gpu_ctx, frames, p)
        monkeypatch.setenv("CVVP_HL_VARIANT", other)
        forced = _gpu(gpu_ctx, frames, p)
        assert np.array_equal(auto, forced), cfg
        for i in (0, 3, n - 1):
            assert np.array_equal(auto[i], ho.highlight_objects(frames[i].copy(), p)), f"{cfg} frame {i} vs oracle"


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


# Openings over structuring elements of many shapes: the compile-time instantiations (4x4 / 5x5 ellipse, 3x3 cross and
# rectangle: csrc/highlight_fused.cu open_bands_ct), the separable run-time plan with one, two and three column patterns,
# the generic tap loop, empty rows, taps without the pixel itself, asymmetric anchors; both kernel builds.
@pytest.mark.parametrize("case", [c for c in hl_cases.adversarial_cases()], ids=lambda c: c[0])
def test_run_time_plan_adversarial_frames_match_oracle(gpu_ctx, case, monkeypatch):
    """CVVP_HL_NO_CT=1: elements that have a compile-time instantiation go through the run-time plan instead"""
    monkeypatch.setenv("CVVP_HL_NO_CT", "1")
    _, frame, p = case
    got = _gpu(gpu_ctx, frame[None], p)[0]
    assert np.array_equal(got, ho.highlight_objects(frame.copy(), p))


@pytest.mark.parametrize("variant", ["large", "small"])
@pytest.mark.parametrize("selem", [
    [[1]], [[1, 1, 1, 1]], [[1], [1], [1], [1]], [[0, 1, 0], [1, 1, 1], [0, 1, 0]], [[1, 1], [1, 1]],
    [[0, 0, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]], [[1, 0, 0, 0, 0, 0, 0, 1]], [[0, 0, 0, 1], [0, 0, 0, 1]],
    [[1, 1, 1], [0, 0, 0], [1, 1, 1]], [[0, 0, 0], [0, 0, 0], [1, 0, 1]], [[1, 0, 1, 1, 0, 1, 1, 1, 1]],
    [[0, 1, 0], [0, 0, 0], [0, 0, 0], [1, 1, 1]], [[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]],
    [[1, 1, 1], [1, 1, 1], [1, 1, 1]], [[0, 0, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [0, 0, 1, 0, 0]],
], ids=lambda s: "x".join(str(len(r)) for r in s[:1]) + f"_{len(s)}rows_" + "".join(str(v) for r in s for v in r)[:12])
def test_openings_of_many_elements_match_oracle(gpu_ctx, selem, variant, monkeypatch):
    """structuring elements of 1-4 rows, taps with and without the pixel itself, empty rows, asymmetric anchors, on frames
    with objects at every border, odd widths included; both kernel builds"""
    monkeypatch.setenv("CVVP_HL_VARIANT", variant)
    rng = np.random.default_rng(len(selem) * 131 + len(selem[0]))
    for (h, w) in ((37, 70), (64, 128), (33, 257), (96, 1000)):
        bg = np.full((h, w), 150, np.uint8)
        frame = bg.copy()
        blobs = rng.integers(0, 2, (h // 4 + 1, w // 4 + 1), dtype=np.uint8).repeat(4, 0).repeat(4, 1)[:h, :w]
        frame[blobs > 0] = 100
        frame[rng.integers(0, h, 40), rng.integers(0, w, 40)] = 90  # specks the opening removes
        frame[:3, :] = 100  # objects on every border
        frame[-2:, :] = 100
        frame[:, :2] = 100
        frame[:, -5:] = 100
        p = ho.HighlightParams(bg, np.array(selem, np.uint8), 14, 7, 16, 3, 3, 0)
        got = _gpu(gpu_ctx, frame[None], p)[0]
        want = ho.highlight_objects(frame.copy(), p)
        assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ at {h}x{w}"
