"""TEST INFRASTRUCTURE ONLY -- loader of oracle/_ref/cvvp_background_ref*.so: the reference's own GetVideoBackground
entry point with everything behind it (Sources/cv_vid_bg_helpers.cpp, Sources/Utility/cv_util.cpp, the AsyncTokens
generator / worker threads, CvVidFramesGeneratorAlgo, CvVidFragmentConsumer, HistogramMedianAlgo8/16/32), compiled
UNMODIFIED from /root/reference against oracle/shim_cv2 (oracle/background_ref_driver.cpp, Makefile target ref_background).

Only tests/ import this module.  It pins what no single class does: the crop rule and its quirk, frame_limit, the
bin-width dispatch, the per-generator frame ranges, the strip split and re-assembly, batch sizes from max_threads."""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mod = None


def path() -> Path:
    return _REF_DIR / ("cvvp_background_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def available() -> bool:
    return path().exists()


def load():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{path()} is missing: run `make -C oracle ref_background` where /root/reference is mounted")
        spec = importlib.util.spec_from_file_location("cvvp_background_ref", path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def get_video_background(vid_path, **pack):
    """GetVideoBackground(VidBgPack{vid_path, bg_algo="hist", max_threads=-1, frame_limit=-1, grayscale, vid_is_grayscale,
    crop_x, crop_y, crop_width, crop_height, token_storage_limit=-1, print_timing_report}) -> ndarray or None"""
    return load().GetVideoBackground(str(vid_path), **pack)
