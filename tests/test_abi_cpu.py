"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/cvvp.h declares, and
fails loudly (no CPU fallback) when no GPU is present.  No compute calls here."""
import ctypes

import pytest

from cvvidproc_b200 import _cabi


def test_header_declares_expected_entry_points():
    names = _cabi.declared_symbols()
    for must in ("cvvp_ctx_create", "cvvp_median_begin", "cvvp_median_push", "cvvp_median_finish",
                 "cvvp_median_device", "cvvp_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    missing = [n for n in _cabi.declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in include/cvvp.h but not exported: {missing}"


def test_ctypes_table_covers_header():
    lib = _cabi.load()
    assert lib.cvvp_abi_version() == 2
    assert set(_cabi.declared_symbols()) <= set(_cabi.BOUND_SYMBOLS)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.CvvpError) as ei:
        _cabi.Context(0)
    assert "no CPU fallback" in str(ei.value) or ei.value.code == -2


def test_null_context_calls_are_rejected():
    lib = _cabi.load()
    assert lib.cvvp_median_begin(None, 10, 1) == -1
    assert lib.cvvp_median_push(None, None, 1, 10) == -1
    assert lib.cvvp_median_finish(None, None) == -1
    assert b"null context" in lib.cvvp_last_error(None)
    assert lib.cvvp_ctx_sm_count(None) == 0
    lib.cvvp_ctx_destroy(None)  # no-op
