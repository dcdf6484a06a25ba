"""CPU tests against the reference's own GetVideoBackground ENTRY POINT, compiled unmodified with everything behind it
(oracle/_ref/cvvp_background_ref, see oracle/background_ref.py): the oracles of the frame source and of the median,
chained, must give what the reference's pipeline gives for the same VidBgPack -- crop rule and quirk, frame_limit,
bin-width dispatch (8 / 16-bit histograms), strip split over any number of workers, the three channel modes.

The reference's calls run in a child process with a progress watchdog (background_ref.run_isolated): a stall of its
thread pipeline must cost a retry, not the test run."""
import numpy as np
import pytest

import report_format
import video_util
from oracle import background_ref as bgref
from oracle import frames_oracle as fo

pytestmark = pytest.mark.skipif(not bgref.available(), reason="oracle/_ref/cvvp_background_ref was not built (no /root/reference)")


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
    jobs, wants = [], []
    for crop in ((0, 0, 0, 0), (3, 5, 33, 20), (69, 47, 0, 0), (10, 0, 500, 0)):
        c = fo.get_cropped_frame_dims(*crop, 70, 48)
        for max_threads, frame_limit in ((1, -1), (2, 25), (5, -1), (8, 10_000), (3, 1)):
            if c[2] < 8 and max_threads > 2:
                continue  # a crop narrower than the strip count makes the reference's own strip split fail (cv_util.cpp:52-56)
            jobs.append((path, dict(max_threads=max_threads, frame_limit=frame_limit, crop_x=crop[0], crop_y=crop[1],
                                    crop_width=crop[2], crop_height=crop[3], **kw)))
            n = len(frames) if frame_limit <= 0 else min(frame_limit, len(frames))
            wants.append(_want(frames, oracle_median, n, c, mode))
    for (_, pack), (bg, _, _), want in zip(jobs, bgref.run_isolated(jobs), wants):
        assert bg is not None and bg.dtype == np.uint8 and bg.shape == want.shape, pack
        assert np.array_equal(bg, want), pack


def test_token_storage_limits_do_not_change_the_result(video, oracle_median):
    """token_storage_limit bounds the reference's queues (token_queue.h:209-214): back-pressure, same image.  (Limits
    below three -- or below the generator count -- stall the reference's pipeline in this build, whatever the cause; they
    are not compared.)"""
    path, frames = video
    want = _want(frames, oracle_median, 60, (0, 0, 70, 48), fo.RGB2GRAY)
    jobs = [(path, dict(max_threads=3, grayscale=True, token_storage_limit=limit)) for limit in (3, 5, 10, -1)]
    for (_, pack), (bg, _, _) in zip(jobs, bgref.run_isolated(jobs)):
        assert np.array_equal(bg, want), pack


def test_bin_width_dispatch_beyond_255_frames(tmp_path, oracle_median):
    """more than 255 frames: the reference switches to 16-bit histograms (cv_vid_bg_helpers.cpp:232-253); with
    frame_limit <= 255 on the same video it uses 8-bit ones.  A pixel that is constant over 300 frames would saturate an
    8-bit counter -- the dispatch is what keeps the result right."""
    frames = _stream(300, 20, 26, 9)
    frames[:, 0, :4] = 200  # constant pixels
    path = video_util.write_lossless(tmp_path / "long.avi", frames)
    assert np.array_equal(video_util.read_all(path), frames)
    limits = (-1, 255, 256)
    jobs = [(path, dict(max_threads=4, frame_limit=limit, grayscale=True)) for limit in limits]
    for limit, (bg, _, _) in zip(limits, bgref.run_isolated(jobs)):
        n = 300 if limit <= 0 else limit
        assert np.array_equal(bg, _want(frames, oracle_median, n, (0, 0, 26, 20), fo.RGB2GRAY)), limit


def test_crop_rule_and_its_quirk_on_a_portrait_video(tmp_path, oracle_median):
    """GetCroppedFrameDims (:39-60) compares height + y against the WIDTH (:56): on a portrait frame a legal height is
    clamped as soon as it exceeds the width; followed by the restatement and by the drop-in module"""
    ref = bgref.load().GetCroppedFrameDims  # plain integer logic, no threads: called in this process
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
    (bg, _, _), = bgref.run_isolated([(path, dict(max_threads=2, vid_is_grayscale=True, crop_x=0, crop_y=30, crop_width=20,
                                                  crop_height=25))])
    assert bg.shape == (30, 20)
    assert np.array_equal(bg, _want(frames, oracle_median, 21, c, fo.CHANNEL0))


def test_entry_point_reports_and_failures(video):
    path, frames = video
    missing, unknown, cropped, timed = bgref.run_isolated([
        ("/no/such/video.avi", {}),
        (path, dict(bg_algo="mean")),  # unknown algorithm (:20-31, :262-266)
        (path, dict(max_threads=2, crop_x=3, crop_y=5, crop_width=33, crop_height=20)),
        (path, dict(max_threads=3, vid_is_grayscale=True, print_timing_report=True)),
    ])
    assert missing[0] is None and "Video file not detected" in missing[2]
    assert unknown[0] is None and "Unknown background algorithm detected: mean" in unknown[2]
    assert "Frames: 60; Res: 70x48(33x20 cropped); FPS: 30" in cropped[1]
    # print_timing_report: what the reference really prints (async_token_process.h:273-414); the drop-in module's report
    # is held to the same expressions on the GPU (tests/test_python_api_gpu.py)
    report_format.check(timed[1])
    assert "(60 batches;" in timed[1] and "(60 tokens;" in timed[1]
