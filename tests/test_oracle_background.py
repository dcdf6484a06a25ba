"""CPU tests against the reference's own GetVideoBackground ENTRY POINT, compiled unmodified with everything behind it
(oracle/_ref/cvvp_background_ref, see oracle/background_ref.py): the oracles of the frame source and of the median,
chained, must give what the reference's pipeline gives for the same VidBgPack -- crop rule and quirk, frame_limit,
bin-width dispatch (8 / 16-bit histograms), strip split over any number of workers, the three channel modes."""
import numpy as np
import pytest

import report_format
import video_util
from oracle import background_ref as bgref
from oracle import frames_oracle as fo

# the reference runs its own threads: a stalled pipeline must end the run, not hang it
pytestmark = [pytest.mark.skipif(not bgref.available(), reason="oracle/_ref/cvvp_background_ref was not built (no /root/reference)"),
              pytest.mark.timeout(180, method="thread")]


def _stream(n, h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.int16)
    return np.clip(base[None] + rng.integers(-20, 21, (n, h, w, 3)), 0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def video(tmp_path_factory):
    frames = _stream(60, 48, 70, 5)
    path = video_util.write_lossless(tmp_path_factory.mktemp("v") / "landscape.avi", frames)
    assert np.array_equal(video_util.read_all(path), frames)
    return path, frames


def _want(frames, oracle_median, n, crop, mode):
    return oracle_median(fo.prepare_frames(frames[:n], crop, mode))


@pytest.mark.parametrize("kw,mode", [({}, fo.AS_IS), ({"grayscale": True}, fo.RGB2GRAY), ({"vid_is_grayscale": True}, fo.CHANNEL0),
                                     ({"grayscale": True, "vid_is_grayscale": True}, fo.CHANNEL0)])
def test_entry_point_matches_the_chained_oracles(video, oracle_median, kw, mode):
    path, frames = video
    for crop in ((0, 0, 0, 0), (3, 5, 33, 20), (69, 47, 0, 0), (10, 0, 500, 0)):
        c = fo.get_cropped_frame_dims(*crop, 70, 48)
        for max_threads, frame_limit in ((1, -1), (2, 25), (5, -1), (8, 10_000), (3, 1)):
            if c[2] < 8 and max_threads > 2:
                continue  # a crop narrower than the strip count makes the reference's own strip split fail (cv_util.cpp:52-56)
            bg = bgref.get_video_background(path, max_threads=max_threads, frame_limit=frame_limit, crop_x=crop[0], crop_y=crop[1],
                                            crop_width=crop[2], crop_height=crop[3], **kw)
            n = len(frames) if frame_limit <= 0 else min(frame_limit, len(frames))
            want = _want(frames, oracle_median, n, c, mode)
            assert bg is not None and bg.dtype == np.uint8 and bg.shape == want.shape, (crop, max_threads, frame_limit)
            assert np.array_equal(bg, want), (crop, max_threads, frame_limit)


def test_token_storage_limits_do_not_change_the_result(video, oracle_median):
    """token_storage_limit bounds the reference's queues (token_queue.h:209-214): back-pressure, same image.  (Limits
    below three -- or below the generator count -- stall the reference's pipeline in this build, whatever the cause; they
    are not compared.)"""
    path, frames = video
    want = _want(frames, oracle_median, 60, (0, 0, 70, 48), fo.RGB2GRAY)
    for limit in (3, 5, 10, -1):
        bg = bgref.get_video_background(path, max_threads=3, grayscale=True, token_storage_limit=limit)
        assert np.array_equal(bg, want), limit


def test_bin_width_dispatch_beyond_255_frames(tmp_path, oracle_median):
    """more than 255 frames: the reference switches to 16-bit histograms (cv_vid_bg_helpers.cpp:232-253); with
    frame_limit <= 255 on the same video it uses 8-bit ones.  A pixel that is constant over 300 frames would saturate an
    8-bit counter -- the dispatch is what keeps the result right."""
    frames = _stream(300, 20, 26, 9)
    frames[:, 0, :4] = 200  # constant pixels
    path = video_util.write_lossless(tmp_path / "long.avi", frames)
    assert np.array_equal(video_util.read_all(path), frames)
    for limit in (-1, 255, 256):
        n = 300 if limit <= 0 else limit
        bg = bgref.get_video_background(path, max_threads=4, frame_limit=limit, grayscale=True)
        assert np.array_equal(bg, _want(frames, oracle_median, n, (0, 0, 26, 20), fo.RGB2GRAY)), limit


def test_crop_rule_and_its_quirk_on_a_portrait_video(tmp_path, oracle_median):
    """GetCroppedFrameDims (:39-60) compares height + y against the WIDTH (:56): on a portrait frame a legal height is
    clamped as soon as it exceeds the width; followed by the restatement and by the drop-in module"""
    ref = bgref.load().GetCroppedFrameDims
    for args in ((0, 0, 0, 0, 640, 480), (10, 20, 100, 50, 640, 480), (600, 0, 100, 0, 640, 480), (0, 200, 0, 300, 640, 480),
                 (0, 100, 0, 400, 480, 640), (5, 7, 1, 1, 9, 11), (0, 30, 20, 25, 40, 60), (39, 59, 5, 5, 40, 60)):
        assert tuple(ref(*args)) == fo.get_cropped_frame_dims(*args), args
    for bad in ((640, 0, 0, 0, 640, 480), (0, 480, 0, 0, 640, 480), (-1, 0, 0, 0, 640, 480), (0, 0, -5, 0, 640, 480)):
        with pytest.raises(RuntimeError):
            ref(*bad)
        with pytest.raises(AssertionError):
            fo.get_cropped_frame_dims(*bad)
    frames = _stream(21, 60, 40, 11)  # portrait: 40 wide, 60 high
    path = video_util.write_lossless(tmp_path / "portrait.avi", frames)
    assert np.array_equal(video_util.read_all(path), frames)
    # y = 30, height = 25: 30 + 25 = 55 > 40 (the width) -> clamped to 60 - 30 = 30 rows although 25 would fit
    crop = (0, 30, 20, 25)
    c = fo.get_cropped_frame_dims(*crop, 40, 60)
    assert c == (0, 30, 20, 30)
    bg = bgref.get_video_background(path, max_threads=2, vid_is_grayscale=True, crop_x=0, crop_y=30, crop_width=20, crop_height=25)
    assert bg.shape == (30, 20)
    assert np.array_equal(bg, _want(frames, oracle_median, 21, c, fo.CHANNEL0))


def test_entry_point_reports_and_failures(video, capfd):
    path, frames = video
    assert bgref.get_video_background("/no/such/video.avi") is None
    assert "Video file not detected" in capfd.readouterr().err
    assert bgref.get_video_background(path, bg_algo="mean") is None  # unknown algorithm (:20-31, :262-266)
    bgref.get_video_background(path, max_threads=2, crop_x=3, crop_y=5, crop_width=33, crop_height=20)
    assert "Frames: 60; Res: 70x48(33x20 cropped); FPS: 30" in capfd.readouterr().out


def test_timing_report_lines_of_the_reference(video, capfd):
    """print_timing_report: what the reference really prints (async_token_process.h:273-414); the drop-in module's report
    is held to the same expressions on the GPU (tests/test_python_api_gpu.py)"""
    path, frames = video
    capfd.readouterr()
    bgref.get_video_background(path, max_threads=3, vid_is_grayscale=True, print_timing_report=True)
    out = capfd.readouterr().out
    report_format.check(out)
    assert "(60 batches;" in out and "(60 tokens;" in out
