"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the per-frame highlight stage.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product path (cvvidproc_b200/csrc, libcvvp_cuda.so) never imports or calls it.

A line-by-line restatement of the reference's HighlightObjectsAlgo
    /root/reference/Sources/ProcessorAlgos/highlight_objects_algo.cpp
        HighlightObjects              :17-78
        ThresholdImage                :81-104
        ThresholdImageWithHysteresis  :107-144
        RemoveSmallObjects            :146-181
        FillHoles                     :183-221
The arithmetic of that file lives in OpenCV (core + imgproc), a third-party dependency that is NOT vendored under
/root/reference (CMakeLists.txt:52 only states `OpenCV >= 4.2.0`; no lock file).  OpenCV C++ is not installed in
this image, but the Python wheel `cv2` 4.13.0 exposes every function the reference calls with the same semantics,
and the reference source itself carries the original Python lines as comments (:26, :34, :38, :116-140, :159-173,
:197-217), so each statement below is the cv2 call of the cited reference line.

PARITY PINNED BY THE REFERENCE'S OWN SOURCE RUN HERE.  The reference ships no tests, golden vectors or fixtures for
this path, so the pin is the reference itself: oracle/Makefile (target ref_highlight) compiles
highlight_objects_algo.{h,cpp} UNMODIFIED from /root/reference against oracle/shim_cv2 -- a stand-in for
<opencv2/opencv.hpp> whose cv:: functions forward to the cv2 wheel, i.e. to OpenCV's real core/imgproc code -- into
oracle/_ref/cvvp_highlight_ref (oracle/highlight_ref_driver.cpp drives the class through Insert / HasResults /
TryGetResult).  tests/test_oracle_highlight.py holds every function below to the reference function it restates
(whole pipeline on 120 random + every adversarial case + full-size C3/C4 frames; ThresholdImage incl. Otsu,
ThresholdImageWithHysteresis, RemoveSmallObjects, FillHoles one by one), and tests/golden/highlight_golden.json holds
hashes of the REFERENCE's outputs (tests/golden/make_highlight_golden.py).  What the pin cannot cover is an OpenCV
other than the wheel's 4.13.0.  The independent label-based model oracle/highlight_model.py stays as a second opinion.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np


@dataclass
class HighlightParams:
    """TokenProcessorPack<HighlightObjectsAlgo> (highlight_objects_algo.h:21-32)."""

    background: np.ndarray
    struct_element: np.ndarray
    threshold: int
    threshold_lo: int
    threshold_hi: int
    min_size_hyst: int
    min_size_threshold: int
    width_border: int = 0  # accepted and unused: the FrameAndFill call is commented out (:68-71)


def threshold_image(image: np.ndarray, threshold: int) -> np.ndarray:
    """ThresholdImage :81-104"""
    if threshold == -1:
        # :94  cv::threshold(image, out, 0, 255, THRESH_BINARY + THRESH_OTSU)
        _, out = cv2.threshold(image, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    else:
        # :100 cv::threshold(image, out, threshold, 255, THRESH_BINARY)
        _, out = cv2.threshold(image, threshold, 255, cv2.THRESH_BINARY)
    return out


def threshold_image_with_hysteresis(image: np.ndarray, threshold_lo: int, threshold_hi: int) -> np.ndarray:
    """ThresholdImageWithHysteresis :107-144"""
    _, thresh_upper = cv2.threshold(image, threshold_hi, 128, cv2.THRESH_BINARY)  # :118
    _, thresh_lower = cv2.threshold(image, threshold_lo, 128, cv2.THRESH_BINARY)  # :122
    # :127 findContours(thresh_upper, RETR_EXTERNAL, CHAIN_APPROX_NONE)
    contours, _ = cv2.findContours(thresh_upper, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    for cnt in contours:
        # :136 floodFill(thresh_lower, contour[0], Scalar(255), 0, 2, 2, FLOODFILL_FIXED_RANGE)  (4-connectivity default)
        seed = (int(cnt[0][0][0]), int(cnt[0][0][1]))
        cv2.floodFill(thresh_lower, None, seed, 255, 2, 2, cv2.FLOODFILL_FIXED_RANGE)
    _, thresh_lower = cv2.threshold(thresh_lower, 200, 255, cv2.THRESH_BINARY)  # :141
    return thresh_lower


def remove_small_objects(image: np.ndarray, min_size_threshold: int) -> None:
    """RemoveSmallObjects :146-181 (in place)"""
    # :161 findContours(image, RETR_TREE, CHAIN_APPROX_SIMPLE)
    contours, _ = cv2.findContours(image, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    small = [c for c in contours if cv2.contourArea(c) < min_size_threshold]  # :169-175 (strict <)
    # :178 drawContours(image, small, -1, Scalar(0,0,0), -1): ALL small contours in one filled draw
    if small:
        cv2.drawContours(image, small, -1, (0, 0, 0), -1)
    # (an empty list is a no-op in C++; cv2's Python binding rejects it, hence the guard)


def fill_holes(image: np.ndarray) -> None:
    """FillHoles :183-221 (in place)"""
    im_floodfill = image.copy()  # :194
    # :201-209  seed = (0,0) if pixel (0,0) == 255, else the bottom-right corner -- follow the code, not its comment
    if im_floodfill.flat[0] == 255:
        pt = (0, 0)
    else:
        pt = (im_floodfill.shape[1] - 1, im_floodfill.shape[0] - 1)
    cv2.floodFill(im_floodfill, None, pt, 255)  # :210 (4-connectivity, exact match)
    im_floodfill = cv2.bitwise_not(im_floodfill)  # :214
    cv2.bitwise_or(image, im_floodfill, dst=image)  # :218


def highlight_objects(frame: np.ndarray, p: HighlightParams, stages: dict | None = None) -> np.ndarray:
    """HighlightObjects :17-78.  Returns the 0/255 mask the reference writes back into the frame token."""
    assert frame.dtype == np.uint8 and frame.ndim == 2
    # :27-29  `background - frame` on two CV_8U Mats is a MatExpr evaluated as cv::subtract -> SATURATING u8
    # (the CV_16S pre-allocation is re-allocated as CV_8U and convertTo(CV_8U) is a no-op); cv::absdiff is commented out
    im_diff = cv2.subtract(p.background, frame)
    selem = np.ascontiguousarray(p.struct_element)

    thresh_bw_1 = threshold_image(im_diff, p.threshold)  # :35
    a0 = thresh_bw_1.copy()
    thresh_bw_1 = cv2.morphologyEx(thresh_bw_1, cv2.MORPH_OPEN, selem)  # :39
    a1 = thresh_bw_1.copy()
    remove_small_objects(thresh_bw_1, p.min_size_threshold)  # :43
    a2 = thresh_bw_1.copy()
    fill_holes(thresh_bw_1)  # :47

    thresh_bw_2 = threshold_image_with_hysteresis(im_diff, p.threshold_lo, p.threshold_hi)  # :54
    b0 = thresh_bw_2.copy()
    thresh_bw_2 = cv2.morphologyEx(thresh_bw_2, cv2.MORPH_OPEN, selem)  # :61
    b1 = thresh_bw_2.copy()
    remove_small_objects(thresh_bw_2, p.min_size_hyst)  # :65
    b2 = thresh_bw_2.copy()
    fill_holes(thresh_bw_2)  # :73 (FrameAndFill at :71 is commented out)

    out = cv2.bitwise_or(thresh_bw_1, thresh_bw_2)  # :77
    if stages is not None:
        stages.update(diff=im_diff, a_thresh=a0, a_open=a1, a_rso=a2, a_fill=thresh_bw_1, b_hyst=b0, b_open=b1,
                      b_rso=b2, b_fill=thresh_bw_2)
    return out


def canonical_struct_element() -> np.ndarray:
    """cv2.getStructuringElement(MORPH_ELLIPSE, (4,4)) as produced by cv2 4.13 (rand_tests.cpp:42 uses that call);
    passed explicitly so that nothing depends on the OpenCV version's ellipse rasteriser."""
    return np.array([[0, 0, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]], np.uint8)


def canonical_params(background: np.ndarray) -> HighlightParams:
    """The only parameter set the reference ever uses (rand_tests.cpp:42-51, :333-342)."""
    return HighlightParams(background=background, struct_element=canonical_struct_element(), threshold=14,
                           threshold_lo=7, threshold_hi=16, min_size_hyst=20, min_size_threshold=20, width_border=5)
