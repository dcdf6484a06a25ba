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
    assert lib.cvvp_abi_version() == 3
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


def test_header_is_plain_c_and_links(tmp_path):
    """the boundary is a C ABI: include/cvvp.h must compile as strict C99 and a C program must link against the library
    (what a cgo / JNI / ctypes-free binding would do); no compute calls"""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "t_abi.c"
    src.write_text(
        '#include "cvvp.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "    cvvp_frame_format f = {64, 48, 3, 0, 0, 64, 48, CVVP_FRAMES_RGB2GRAY};\n"
        '    printf("%d %zu %zu %d\\n", cvvp_abi_version(), cvvp_frame_format_out_bytes(&f), sizeof(cvvp_component),\n'
        "           cvvp_highlight_queue_pending(NULL));\n"
        "    return cvvp_median_begin(NULL, 1, 1) == CVVP_ERR_INVALID ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "t_abi"
    lib_dir = _cabi.LIB_PATH.parent
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(_cabi.HEADER.parent), str(src),
                    "-L", str(lib_dir), "-lcvvp_cuda", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["3", "3072", "48", "0"]
