"""TEST INFRASTRUCTURE ONLY -- loader of oracle/_ref/cvvp_frames_ref*.so: the reference's own CvVidFramesGeneratorAlgo
(/root/reference/Sources/ProcessorTokenHandlers/cv_vid_frames_generator_algo.h) compiled UNMODIFIED against
oracle/shim_cv2 (oracle/frames_ref_driver.cpp, oracle/Makefile target ref_frames).

Only tests/ and tests/golden/make_frames_golden.py import this module.  It pins oracle/frames_oracle.py: the tokens
the reference's generator emits for a lossless video (frame range, crop, the three channel modes) are what the
restatement computes from the decoded frames.
"""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mod = None


def path() -> Path:
    return _REF_DIR / ("cvvp_frames_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def available() -> bool:
    return path().exists()


def load():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{path()} is missing: run `make -C oracle ref_frames` where /root/reference is mounted")
        spec = importlib.util.spec_from_file_location("cvvp_frames_ref", path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def tokens(vid_path, start_frame, last_frame, crop, mode, frames_in_batch=4):
    """Every token of the generator, batch after batch until GetTokenSet() returns an empty set.
    mode: frames_oracle.AS_IS / CHANNEL0 (vid_is_grayscale) / RGB2GRAY (convert_to_grayscale)."""
    x, y, w, h = crop
    gen = load().RefFrames(str(vid_path), start_frame, last_frame, x, y, w, h, mode == 2, mode == 1, frames_in_batch)
    out = []
    while True:
        batch = gen.get_token_set()
        if not batch:
            return out
        out += batch
