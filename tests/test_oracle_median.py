"""CPU tests: the median oracle (oracle/median_oracle.c) against the golden vectors produced by the reference's
own class (tests/golden/make_golden.py), against the reference class itself where oracle/_ref exists, and against
numpy's order statistic.  Also pins the host synthetic generator through the input hashes."""
import hashlib
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN_DIR = Path(__file__).parent / "golden"
GOLDEN = json.loads((GOLDEN_DIR / "median_golden.json").read_text())

_spec = importlib.util.spec_from_file_location("make_golden", GOLDEN_DIR / "make_golden.py")
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_oracle_reproduces_reference_golden(case, oracle_median):
    frames = make_golden.case_input(case)
    assert hashlib.sha256(frames.tobytes()).hexdigest() == case["input_sha256"], "input generator drifted"
    got = oracle_median(frames, bin_bytes=case.get("bin_bytes", 0), nthreads=3)
    assert hashlib.sha256(got.tobytes()).hexdigest() == case["output_sha256"]
    if "output" in case:
        assert got.reshape(-1).tolist() == case["output"]
    if not case.get("bin_bytes"):
        # without counter saturation the reference's rule is exactly sorted[N/2] (histogram_median_algo.h:160-166)
        assert np.array_equal(got, np.sort(frames, axis=0)[frames.shape[0] // 2])


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_reference_build_reproduces_golden(case, ref_median):
    if ref_median is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    frames = make_golden.case_input(case)
    got = ref_median(frames, bin_bytes=case.get("bin_bytes", 0), nthreads=2)
    assert hashlib.sha256(got.tobytes()).hexdigest() == case["output_sha256"]


def test_saturation_backtrack_matches_reference(oracle_median, ref_median):
    """histogram_median_algo.h:169-184 -- only reachable when the counters saturate; oracle == reference class."""
    if ref_median is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(42)
    for n in (256, 300, 511, 700):
        frames = rng.integers(0, 3, (n, 6, 7), dtype=np.uint8)  # 3 values -> counts exceed 255
        assert np.array_equal(oracle_median(frames, bin_bytes=1, nthreads=1), ref_median(frames, bin_bytes=1, nthreads=1))


def test_strip_count_does_not_change_result(oracle_median):
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (64, 10, 21), dtype=np.uint8)
    base = oracle_median(frames, nthreads=1)
    for t in (2, 3, 7, 16):
        assert np.array_equal(oracle_median(frames, nthreads=t), base)


def test_bin_width_rule(oracle_lib):
    import ctypes

    f = oracle_lib.cvvp_oracle_bin_bytes_for
    f.argtypes = [ctypes.c_longlong]
    assert [f(n) for n in (1, 255, 256, 65535, 65536, 2**32 - 1, 2**32)] == [1, 1, 2, 2, 4, 4, 0]
