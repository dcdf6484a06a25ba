"""TEST INFRASTRUCTURE ONLY -- loader of oracle/_ref/cvvp_binding_ref*.so: INTEGRATION.md's reference-side bindings
(integration/gpu_median_algo.h, integration/gpu_highlight_algo.h) compiled against the reference's own
token_processor_algo.h and the cv2-forwarding cv::Mat (oracle/binding_ref_driver.cpp, oracle/Makefile target
ref_binding), linked with libcvvp_cuda.so.  Not an oracle: it computes nothing itself; tests use it to drive the
product through the reference's plugin interface."""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path

_REF_DIR = Path(__file__).resolve().parent / "_ref"
_mod = None


def path() -> Path:
    return _REF_DIR / ("cvvp_binding_ref" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def available() -> bool:
    return path().exists()


def load():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(f"{path()} is missing: run `make -C oracle ref_binding` where /root/reference is mounted")
        spec = importlib.util.spec_from_file_location("cvvp_binding_ref", path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
