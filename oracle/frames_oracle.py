"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the frame source (crop + channel reduction of decoded frames).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(cvvidproc_b200/csrc/frames.cu behind cvvp_frames_prepare* / cvvp_median_push_source / cvvp_highlight_queue_*) never does.

A line-by-line restatement of the per-frame work of the reference's generator
    /root/reference/Sources/ProcessorTokenHandlers/cv_vid_frames_generator_algo.h
        frame = frame(m_pack.crop_rectangle)                        :141
        cv::extractChannel(frame, modified_frame, 0)                :149-151   (vid_is_grayscale)
        cv::cvtColor(frame, modified_frame, cv::COLOR_RGB2GRAY)     :152-154   (convert_to_grayscale)
        modified_frame = std::move(frame)                           :155-156   (neither)
and of the crop-rectangle rule
    /root/reference/Sources/cv_vid_bg_helpers.cpp  GetCroppedFrameDims :39-60  (incl. the height-vs-hor_pixels quirk :56).
The arithmetic of cvtColor lives in OpenCV, which is NOT vendored under /root/reference (CMakeLists.txt:52 states
`OpenCV >= 4.2.0`); each statement below is the cv2 4.13.0 call of the cited line.

PARITY PINNED BY THE REFERENCE'S OWN SOURCE RUN HERE.  The reference has no tests or fixtures for this path
(SURVEY.md section 4); the pin is its generator itself: oracle/Makefile (target ref_frames) compiles
cv_vid_frames_generator_algo.h UNMODIFIED from /root/reference against oracle/shim_cv2 (cv::VideoCapture,
cv::extractChannel, cv::cvtColor forward to the cv2 wheel) into oracle/_ref/cvvp_frames_ref, and
tests/test_oracle_frames.py holds `prepare_frames` to the tokens GetTokenSet() emits for lossless videos (frame range,
crop, all three channel modes); tests/golden/frames_golden.json holds hashes of the reference's tokens wherever a video
can carry the case (tests/golden/make_frames_golden.py).  GetCroppedFrameDims is held to the reference's own function, and the crop +
frame range + channel mode chain to the reference's GetVideoBackground entry point (oracle/_ref/cvvp_background_ref,
tests/test_oracle_background.py).  Beyond that:
`rgb2gray_fixed_point` below (the closed form of OpenCV's 8-bit RGB2GRAY) is held to cv2 on ALL 2^24 colour triples
by tests/test_oracle_frames.py.
"""
from __future__ import annotations

import cv2
import numpy as np

AS_IS, CHANNEL0, RGB2GRAY = 0, 1, 2


def get_cropped_frame_dims(x: int, y: int, width: int, height: int, hor_pixels: int, vert_pixels: int):
    """GetCroppedFrameDims, cv_vid_bg_helpers.cpp:39-60 -> (x, y, width, height)"""
    assert x >= 0 and y >= 0 and width >= 0 and height >= 0            # :42-45
    assert hor_pixels > 0 and vert_pixels > 0                          # :46-47
    assert x < hor_pixels and y < vert_pixels                          # :48-49
    if width == 0 or width + x > hor_pixels:                           # :52-53
        width = hor_pixels - x
    if height == 0 or height + y > hor_pixels:                         # :56-57 (compares against hor_pixels: followed, not fixed)
        height = vert_pixels - y
    return x, y, width, height


def prepare_frame(frame: np.ndarray, crop, mode: int) -> np.ndarray:
    """One decoded frame (H, W) or (H, W, C) -> the token the generator emits (:141-156)."""
    x, y, w, h = crop
    f = frame[y:y + h, x:x + w]                                        # :141
    if mode == CHANNEL0:
        out = cv2.extractChannel(np.ascontiguousarray(f), 0) if f.ndim == 3 else f   # :151
    elif mode == RGB2GRAY:
        out = cv2.cvtColor(np.ascontiguousarray(f), cv2.COLOR_RGB2GRAY)               # :154
    else:
        out = f                                                        # :156
    return np.ascontiguousarray(out)


def prepare_frames(frames: np.ndarray, crop, mode: int) -> np.ndarray:
    return np.stack([prepare_frame(f, crop, mode) for f in frames])


def rgb2gray_fixed_point(c0, c1, c2):
    """Closed form of OpenCV's 8-bit COLOR_RGB2GRAY (15-bit coefficients, round to nearest); c0 is the first channel
    in memory.  Held to cv2 on all 2^24 triples by tests/test_oracle_frames.py."""
    c0 = np.asarray(c0, np.int64)
    c1 = np.asarray(c1, np.int64)
    c2 = np.asarray(c2, np.int64)
    return ((c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15).astype(np.uint8)
