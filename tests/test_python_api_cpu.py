"""CPU tests of the drop-in Python surface (cvvidproc_b200._core vs Sources/py_bindings.cpp): names, argument order and
defaults, and the failure modes that return before any device work (missing video, unknown algorithm, validation)."""
import inspect

import numpy as np
import pytest

import cvvidproc_b200 as cvp
import video_util


def test_reexported_names_match_reference_package():
    # PySources/cvvidproc/__init__.py:3
    for name in ("VidBgPack", "GetVideoBackground", "HighlightObjectsPack", "AssignObjectsPack", "VidObjectTrackPack",
                 "TrackObjects"):
        assert hasattr(cvp, name)
    import cvvidproc_b200._core as core

    assert core.__doc__ == "C++ bindings for processing an opencv video"  # py_bindings.cpp:33


def _sig_names(doc):
    head = doc.split("->")[0]
    return [part.split(":")[0].strip() for part in head[head.index("(") + 1 :].split(", ") if ":" in part]


def test_pack_signatures_and_defaults():
    doc = cvp.VidBgPack.__init__.__doc__
    assert _sig_names(doc)[1:] == ["vid_path", "bg_algo", "max_threads", "frame_limit", "grayscale", "vid_is_grayscale",
                                   "crop_x", "crop_y", "crop_width", "crop_height", "token_storage_limit",
                                   "print_timing_report"]  # py_bindings.cpp:36-60
    for frag in ("bg_algo: str = 'hist'", "= -1, frame_limit", "token_storage_limit", "= 10", "print_timing_report: bool = False"):
        assert frag in doc
    doc = cvp.VidObjectTrackPack.__init__.__doc__
    assert _sig_names(doc)[1:] == ["vid_path", "highlight_objects_pack", "assign_objects_pack", "max_threads", "start_frame",
                                   "frame_limit", "grayscale", "vid_is_grayscale", "crop_x", "crop_y", "crop_width",
                                   "crop_height", "token_storage_limit", "print_timing_report"]  # :98-126
    doc = cvp.HighlightObjectsPack.__init__.__doc__
    assert _sig_names(doc)[1:] == ["background", "struct_element", "threshold", "threshold_lo", "threshold_hi",
                                   "min_size_hyst", "min_size_threshold", "width_border"]  # :69-85
    assert _sig_names(cvp.AssignObjectsPack.__init__.__doc__)[1:] == ["function", "kwargs"]  # :88-95
    # keyword construction works with the reference's names
    cvp.VidBgPack(vid_path="x.mp4", bg_algo="hist", max_threads=-1, frame_limit=5, grayscale=True, vid_is_grayscale=False,
                  crop_x=0, crop_y=0, crop_width=0, crop_height=0, token_storage_limit=10, print_timing_report=False)


def test_missing_video_returns_none_and_empty_dict(capfd):
    # cv_vid_bg_helpers.cpp:202-207 / cv_vid_objecttrack_helpers.cpp:158-163: message on stderr, empty result, no raise
    assert cvp.GetVideoBackground(cvp.VidBgPack("/nonexistent/video.mp4")) is None
    hp = cvp.HighlightObjectsPack(np.zeros((4, 4), np.uint8), np.ones((2, 2), np.uint8), 1, 1, 1, 1, 1, 1)
    ap = cvp.AssignObjectsPack(lambda **kw: 0, {})
    assert cvp.TrackObjects(cvp.VidObjectTrackPack("/nonexistent/video.mp4", hp, ap)) == {}
    err = capfd.readouterr().err
    assert err.count("Video file not detected: /nonexistent/video.mp4") == 2


@pytest.fixture(scope="module")
def tiny_video(tmp_path_factory):
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (6, 24, 32), dtype=np.uint8)
    return video_util.write_lossless(tmp_path_factory.mktemp("vid") / "tiny.avi", frames), frames


def test_lossless_fixture_roundtrip(tiny_video):
    path, frames = tiny_video
    back = video_util.read_all(path)
    assert back.shape == (6, 24, 32, 3)
    assert np.array_equal(back[..., 0], frames)


def test_unknown_algorithm_returns_none(tiny_video, capfd):
    path, _ = tiny_video
    assert cvp.GetVideoBackground(cvp.VidBgPack(path, bg_algo="nope")) is None  # cv_vid_bg_helpers.cpp:255-260
    cap = capfd.readouterr()
    assert "tried to get vid background with unknown algorithm: nope" in cap.err
    assert "Frames: 6; Res: 32x24; FPS: 30" in cap.out  # the info line is printed first (:212-223)


def test_trackobjects_validation_raises_runtime_error(tiny_video):
    path, _ = tiny_video
    ap = cvp.AssignObjectsPack(lambda **kw: 0, {})
    se = np.ones((2, 2), np.uint8)
    # background size != cropped frame size (cv_vid_objecttrack_helpers.cpp:171-172)
    hp = cvp.HighlightObjectsPack(np.zeros((10, 10), np.uint8), se, 1, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="assert failed"):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, vid_is_grayscale=True))
    # empty background (:167) and empty structuring element (:175)
    hp = cvp.HighlightObjectsPack(np.zeros((0, 0), np.uint8), se, 1, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, vid_is_grayscale=True))
    hp = cvp.HighlightObjectsPack(np.zeros((24, 32), np.uint8), np.zeros((0, 0), np.uint8), 1, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, vid_is_grayscale=True))
    # crop window starting outside the frame (GetCroppedFrameDims, cv_vid_bg_helpers.cpp:41-49)
    hp = cvp.HighlightObjectsPack(np.zeros((24, 32), np.uint8), se, 1, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="start of crop window"):
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, ap, vid_is_grayscale=True, crop_x=100))
