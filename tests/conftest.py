import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """The native pieces are built in-tree by __graft_entry__.build(); build on demand for a fresh checkout."""
    from cvvidproc_b200 import build

    if not build.LIB_PATH.exists():
        build.build_cuda_lib()
    if not (REPO / "oracle" / "_build" / "libcvvp_oracle.so").exists():
        build.build_oracle()


_ensure_built()

_ORACLE_SIG = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
               ctypes.c_void_p]


def _median_via(lib, fn_name, frames, bin_bytes=0, nthreads=1):
    frames = np.ascontiguousarray(frames)
    n = frames.shape[0]
    nelem = int(np.prod(frames.shape[1:]))
    out = np.empty(nelem, np.uint8)
    fn = getattr(lib, fn_name)
    fn.argtypes = _ORACLE_SIG
    fn.restype = ctypes.c_int
    rc = fn(frames.ctypes.data, n, nelem, nelem, bin_bytes, nthreads, out.ctypes.data)
    assert rc == 0, f"{fn_name} failed with {rc}"
    return out.reshape(frames.shape[1:])


@pytest.fixture(scope="session")
def oracle_lib():
    return ctypes.CDLL(str(REPO / "oracle" / "_build" / "libcvvp_oracle.so"))


@pytest.fixture(scope="session")
def oracle_median(oracle_lib):
    """CPU restatement of HistogramMedianAlgo (oracle/median_oracle.c). Checker only."""

    def run(frames, bin_bytes=0, nthreads=4):
        return _median_via(oracle_lib, "cvvp_oracle_median", frames, bin_bytes, nthreads)

    return run


@pytest.fixture(scope="session")
def ref_median():
    """The reference's own class compiled from /root/reference (oracle/_ref); None where it was never built."""
    path = REPO / "oracle" / "_ref" / "libcvvp_median_ref.so"
    if not path.exists():
        return None
    lib = ctypes.CDLL(str(path))

    def run(frames, bin_bytes=0, nthreads=4):
        return _median_via(lib, "cvvp_ref_median", frames, bin_bytes, nthreads)

    return run


@pytest.fixture(scope="session")
def gpu_ctx():
    from cvvidproc_b200 import _cabi

    ctx = _cabi.Context(0)
    yield ctx
    ctx.close()
