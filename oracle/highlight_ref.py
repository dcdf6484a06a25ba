"""TEST INFRASTRUCTURE ONLY -- loader of oracle/_ref/cvvp_highlight_ref*.so: the reference's own
HighlightObjectsAlgo (/root/reference/Sources/ProcessorAlgos/highlight_objects_algo.{h,cpp}) compiled UNMODIFIED
against oracle/shim_cv2, whose cv:: functions forward to the cv2 wheel (oracle/highlight_ref_driver.cpp,
oracle/Makefile target ref_highlight).

Only tests/ and tests/golden/make_highlight_golden.py import this module.  It is what pins oracle/highlight_oracle.py:
every function of the restatement is held to the reference function it restates (tests/test_oracle_highlight.py),
and the golden hashes of tests/golden/highlight_golden.json are taken from the reference's output.
"""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mod = None


def path() -> Path:
    return _REF_DIR / ("cvvp_highlight_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def available() -> bool:
    return path().exists()


def load():
    """The extension module; raises when it was not built (it is built wherever /root/reference is mounted)."""
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{path()} is missing: run `make -C oracle ref_highlight` where /root/reference is mounted")
        spec = importlib.util.spec_from_file_location("cvvp_highlight_ref", path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def operator(p):
    """HighlightObjectsAlgo{TokenProcessorPack<HighlightObjectsAlgo>{...}} for an oracle.highlight_oracle.HighlightParams"""
    return load().RefHighlight(p.background, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi,
                               p.min_size_hyst, p.min_size_threshold, p.width_border)


def highlight_objects(frame, p):
    """one token through Insert / HasResults / TryGetResult of a fresh operator"""
    return operator(p).insert(frame)
