"""GPU end-to-end tests of the drop-in Python surface: GetVideoBackground / TrackObjects on lossless test videos,
compared with the CPU oracles driven the way the reference drives its algos."""
import cv2
import numpy as np
import pytest

import cvvidproc_b200 as cvp
import report_format
import video_util
from cvvidproc_b200 import synth
from oracle import highlight_oracle as ho

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gray_video(tmp_path_factory):
    frames = synth.synth_frames(0, 60, 160, 96, 7, 8)
    return video_util.write_lossless(tmp_path_factory.mktemp("v") / "gray.avi", frames), frames


@pytest.fixture(scope="module")
def color_video(tmp_path_factory):
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (25, 40, 56, 3), dtype=np.uint8)
    return video_util.write_lossless(tmp_path_factory.mktemp("v") / "color.avi", frames), frames


def test_background_of_gray_video(gray_video, oracle_median, capfd):
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    assert bg.dtype == np.uint8 and bg.shape == (96, 160)
    assert np.array_equal(bg, oracle_median(frames))
    assert "Frames: 60; Res: 160x96; FPS: 30" in capfd.readouterr().out
    # frame_limit (cv_vid_bg_helpers.cpp:226-229) and values beyond the video length
    bg10 = cvp.GetVideoBackground(cvp.VidBgPack(path, frame_limit=10, vid_is_grayscale=True))
    assert np.array_equal(bg10, oracle_median(frames[:10]))
    bg_all = cvp.GetVideoBackground(cvp.VidBgPack(path, frame_limit=10_000, vid_is_grayscale=True))
    assert np.array_equal(bg_all, bg)


def test_background_of_a_video_longer_than_the_device_holds(gray_video, color_video, oracle_median, monkeypatch):
    """GetVideoBackground with the resident stack capped (what a video beyond the device's memory does): frames are
    folded into value histograms as they arrive (csrc/median_hist.cu); same background"""
    monkeypatch.setenv("CVVP_MEDIAN_RESIDENT_MAX", "16")
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    assert np.array_equal(bg, oracle_median(frames))
    path, frames = color_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, crop_x=5, crop_y=3, crop_width=30, crop_height=20))
    assert np.array_equal(bg, oracle_median(np.ascontiguousarray(frames[:, 3:23, 5:35])))


def test_background_crop_and_color_modes(color_video, oracle_median, capfd):
    path, frames = color_video  # BGR as stored
    # no grayscale flag: element-wise median over all three channels, result (H, W, 3) (ndarray_converter.cpp:141-142)
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path))
    assert bg.shape == (40, 56, 3)
    assert np.array_equal(bg, oracle_median(frames))
    # grayscale=True: COLOR_RGB2GRAY applied to the frames as decoded (cv_vid_frames_generator_algo.h:152-154)
    gray = np.stack([cv2.cvtColor(f, cv2.COLOR_RGB2GRAY) for f in frames])
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, grayscale=True))
    assert bg.shape == (40, 56)
    assert np.array_equal(bg, oracle_median(gray))
    # vid_is_grayscale=True: channel 0 (:149-151); with a crop window
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True, crop_x=5, crop_y=3, crop_width=30, crop_height=20))
    assert bg.shape == (20, 30)
    assert np.array_equal(bg, oracle_median(np.ascontiguousarray(frames[:, 3:23, 5:35, 0])))
    assert "(30x20 cropped)" in capfd.readouterr().out


def _tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
    """deterministic stand-in for the external bubble tracker: 8-connected components of the mask"""
    assert bw_frame.dtype == np.uint8 and set(np.unique(bw_frame)) <= {0, 255}
    n, _, stats, cent = cv2.connectedComponentsWithStats(bw_frame, connectivity=8)
    objects_prev.clear()
    for i in range(1, n):
        if stats[i, cv2.CC_STAT_AREA] < kwargs["min_area"]:
            continue
        objects_prev[next_ID] = (frames_processed, int(stats[i, cv2.CC_STAT_AREA]))
        objects_archive[next_ID] = {"frame": frames_processed, "area": int(stats[i, cv2.CC_STAT_AREA]),
                                    "cx": round(float(cent[i][0]), 3), "cy": round(float(cent[i][1]), 3)}
        next_ID += 1
    return next_ID


def _reference_track(frames, p, kwargs):
    """what the reference computes: HighlightObjects per frame, then the callback strictly in order"""
    prev, archive, nid = {}, {}, 0
    for i, f in enumerate(frames):
        bw = ho.highlight_objects(f.copy(), p)
        nid = _tracker(bw, i, prev, archive, nid, kwargs)
    return archive


def test_track_objects_end_to_end(gray_video):
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    p = ho.canonical_params(bg)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    kwargs = {"min_area": 5}
    ap = cvp.AssignObjectsPack(_tracker, kwargs)
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, vid_is_grayscale=True))
    want = _reference_track(frames, p, kwargs)
    assert len(want) > 10
    assert archive == want
    # start_frame / frame_limit window (cv_vid_objecttrack_helpers.cpp:54-80): frames_processed restarts at 0
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, start_frame=20, frame_limit=15, vid_is_grayscale=True))
    assert archive == _reference_track(frames[20:35], p, kwargs)


def _tracker_with_components(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs, components):
    """the same tracker fed by the device's components (kwargs["cvvp_components"] = True) instead of a host CCL; it
    also checks them against cv2 on the mask it was handed"""
    n, lab, stats, cent = cv2.connectedComponentsWithStats(bw_frame, connectivity=8)
    assert components["count"] == n - 1 == len(components["stats"])
    if "labels" in components:
        pairs = np.unique(np.stack([components["labels"].ravel(), lab.ravel()], axis=1), axis=0)
        assert len(pairs[pairs[:, 0] != 0]) == n - 1 and not np.any((pairs[:, 0] == 0) != (pairs[:, 1] == 0))  # same partition
    objects_prev.clear()
    for k in range(components["count"]):
        x, y, w, h, area = (int(v) for v in components["stats"][k])
        if area < kwargs["min_area"]:
            continue
        cx, cy = components["centroids"][k]
        objects_prev[next_ID] = (frames_processed, area)
        objects_archive[next_ID] = {"frame": frames_processed, "area": area, "cx": round(float(cx), 3), "cy": round(float(cy), 3)}
        next_ID += 1
    return next_ID


def test_track_objects_with_device_components(gray_video):
    """opt-in: the device labels the masks; cv2 numbers components in its own scan order, so the archives are compared
    as sets of per-frame records"""
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    p = ho.canonical_params(bg)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    want = _reference_track(frames, p, {"min_area": 5})
    for extra in ({}, {"cvvp_labels": True}):
        kwargs = {"min_area": 5, "cvvp_components": True, **extra}
        archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker_with_components, kwargs),
                                                          vid_is_grayscale=True))
        key = lambda d: sorted((v["frame"], v["area"], v["cx"], v["cy"]) for v in d.values())  # noqa: E731
        assert key(archive) == key(want)


def test_timing_reports_follow_the_reference_format(gray_video, capfd):
    """print_timing_report: the report lines of AsyncTokenProcess::GetTimingInfoAndResetTimer
    (async_token_process.h:273-414) and the two headings of cv_vid_objecttrack_helpers.cpp:136-143"""
    import re

    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True, print_timing_report=True))
    out = capfd.readouterr().out
    report_format.check(out)  # the same expressions match the reference's own output (tests/test_oracle_background.py)
    assert "(60 batches;" in out and "(60 tokens;" in out
    p = ho.canonical_params(bg)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker, {"min_area": 5}), vid_is_grayscale=True,
                                            print_timing_report=True))
    out = capfd.readouterr().out
    assert out.index("Highlight objects timing report:") < out.index("Assign objects timing report:")
    assert len(re.findall(r"^Unit \[1\]: ", out, re.M)) == 2 and "(60 tokens;" in out


def test_track_objects_needs_single_channel(color_video):
    path, frames = color_video
    hp = cvp.HighlightObjectsPack(np.zeros((40, 56), np.uint8), np.ones((2, 2), np.uint8), 1, 1, 1, 1, 1, 1)
    ap = cvp.AssignObjectsPack(lambda **kw: 0, {})
    with pytest.raises(RuntimeError, match="single-channel"):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap))


def test_device_and_host_frame_preparation_agree(color_video, gray_video, oracle_median, monkeypatch):
    """the generator's crop / channel reduction runs on the device by default (csrc/frames.cu); CVVP_HOST_PREP=1 keeps it
    on the host through cv2 -- same backgrounds, same archives"""
    path, frames = color_video
    packs = [dict(), dict(grayscale=True), dict(vid_is_grayscale=True), dict(grayscale=True, crop_x=3, crop_y=2, crop_width=41, crop_height=30),
             dict(crop_x=1, crop_y=1, crop_width=50, crop_height=35)]
    dev = [cvp.GetVideoBackground(cvp.VidBgPack(path, **kw)) for kw in packs]
    monkeypatch.setenv("CVVP_HOST_PREP", "1")
    host = [cvp.GetVideoBackground(cvp.VidBgPack(path, **kw)) for kw in packs]
    monkeypatch.delenv("CVVP_HOST_PREP")
    for d, h in zip(dev, host):
        assert d.shape == h.shape and np.array_equal(d, h)
    gray = np.stack([cv2.cvtColor(np.ascontiguousarray(f[2:32, 3:44]), cv2.COLOR_RGB2GRAY) for f in frames])
    assert np.array_equal(dev[3], oracle_median(gray))


def test_track_objects_pipeline_of_small_batches(gray_video, monkeypatch):
    """several batches in flight (CVVP_TRACK_BATCH=7, token_storage_limit 1..3): the callback still sees every frame once,
    in order, and the archive is the reference's"""
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    p = ho.canonical_params(bg)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    kwargs = {"min_area": 5}
    want = _reference_track(frames, p, kwargs)
    monkeypatch.setenv("CVVP_TRACK_BATCH", "7")
    for limit in (1, 2, 10):
        seen = []

        def tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
            seen.append(frames_processed)
            return _tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs)

        archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(tracker, kwargs), vid_is_grayscale=True,
                                                          token_storage_limit=limit))
        assert seen == list(range(len(frames)))
        assert archive == want
    # the same with the device's components, and with the host-side frame preparation
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker_with_components,
                                                                                     {"min_area": 5, "cvvp_components": True}),
                                                      vid_is_grayscale=True))
    key = lambda d: sorted((v["frame"], v["area"], v["cx"], v["cy"]) for v in d.values())  # noqa: E731
    assert key(archive) == key(want)
    monkeypatch.setenv("CVVP_HOST_PREP", "1")
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker, kwargs), vid_is_grayscale=True))
    assert archive == want


def test_track_objects_on_a_cropped_colour_video(color_video):
    """grayscale=True on a colour video with a crop window: RGB2GRAY of the decoded frames on the device, masks and
    archive equal to the cv2 restatement on cv2-prepared frames"""
    path, frames = color_video
    crop = dict(crop_x=4, crop_y=2, crop_width=48, crop_height=36)
    gray = np.stack([cv2.cvtColor(np.ascontiguousarray(f[2:38, 4:52]), cv2.COLOR_RGB2GRAY) for f in frames])
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, grayscale=True, **crop))
    assert bg.shape == (36, 48)
    p = ho.HighlightParams(bg, np.ones((2, 2), np.uint8), 40, 25, 60, 2, 2, 0)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    kwargs = {"min_area": 1}
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker, kwargs), grayscale=True, **crop))
    want = _reference_track(gray, p, kwargs)
    assert len(want) > 5
    assert archive == want


def _device_lists():
    import torch

    lists = ["0,0", "0,0,0"]  # several operators on one device: the ordering / exchange logic without a second GPU
    if torch.cuda.device_count() >= 2:
        lists.append("0,1")
    return lists


def test_several_devices_one_stream(gray_video, color_video, oracle_median, monkeypatch):
    """CVVP_DEVICES spreads both entry points over several devices: TrackObjects hands batches to the devices in turn
    and still delivers the masks strictly in frame order (mat_set_intermediary.h:50-68,84-114); GetVideoBackground
    gives every device a share of the frames and merges them with the frame-sharded median
    (cv_vid_bg_helpers.cpp:84-120).  Same archive, same background as with one device."""
    path, frames = gray_video
    cpath, cframes = color_video
    bg1 = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    assert np.array_equal(bg1, oracle_median(frames))
    p = ho.canonical_params(bg1)
    hp = cvp.HighlightObjectsPack(bg1, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)
    kwargs = {"min_area": 5}
    want = _reference_track(frames, p, kwargs)
    monkeypatch.setenv("CVVP_TRACK_BATCH", "4")
    for devs in _device_lists():
        monkeypatch.setenv("CVVP_DEVICES", devs)
        bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
        assert np.array_equal(bg, bg1), devs
        bgc = cvp.GetVideoBackground(cvp.VidBgPack(cpath, grayscale=True, crop_x=3, crop_y=2, crop_width=41, crop_height=30))
        gray = np.stack([cv2.cvtColor(np.ascontiguousarray(f[2:32, 3:44]), cv2.COLOR_RGB2GRAY) for f in cframes])
        assert np.array_equal(bgc, oracle_median(gray)), devs
        seen = []

        def tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
            seen.append(frames_processed)
            return _tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs)

        for limit in (1, 3):
            seen.clear()
            archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(tracker, kwargs),
                                                              vid_is_grayscale=True, token_storage_limit=limit))
            assert seen == list(range(len(frames))), (devs, limit)
            assert archive == want, (devs, limit)


def test_other_python_threads_run_during_the_calls(gray_video):
    """The reference releases the GIL in GetVideoBackground (py_bindings.cpp:63-66); here the uploads, the device work
    and the waits run without it (and cv2 drops it while decoding), so a second Python thread keeps making progress --
    also during TrackObjects, whose callbacks are the only part that needs the GIL."""
    import threading
    import time

    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    ticks, stop = [0], threading.Event()

    def spin():
        while not stop.is_set():
            ticks[0] += 1
            time.sleep(0.0005)

    t = threading.Thread(target=spin)
    t.start()
    try:
        before = ticks[0]
        t0 = time.perf_counter()
        for _ in range(3):
            cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
        dt = time.perf_counter() - t0
        during_bg = ticks[0] - before
        p = ho.canonical_params(bg)
        hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                      p.min_size_threshold, p.width_border)
        before = ticks[0]
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker, {"min_area": 5}), vid_is_grayscale=True))
        during_track = ticks[0] - before
    finally:
        stop.set()
        t.join()
    # a thread that sleeps 0.5 ms per tick gets at most ~1000-2000 ticks per second; demand a clear share of that
    assert during_bg >= max(5, int(200 * dt)), (during_bg, dt)
    assert during_track >= 3


def test_a_failing_callback_ends_the_call_cleanly(gray_video):
    """an exception in the tracker callback propagates (the decode thread is stopped, the queues are drained) and the
    module stays usable"""
    path, frames = gray_video
    bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True))
    p = ho.canonical_params(bg)
    hp = cvp.HighlightObjectsPack(bg, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                                  p.min_size_threshold, p.width_border)

    def bad(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
        if frames_processed == 9:
            raise ValueError("tracker gave up")
        return next_ID

    with pytest.raises(ValueError, match="tracker gave up"):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(bad, {}), vid_is_grayscale=True))
    archive = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(_tracker, {"min_area": 5}), vid_is_grayscale=True))
    assert archive == _reference_track(frames, p, {"min_area": 5})
