"""Generates tests/golden/median_golden.json.

Run in the build container, where /root/reference is mounted and oracle/_ref was built from it
(make -C oracle): every entry is the output of the REFERENCE's own HistogramMedianAlgo<T> class
(Sources/ProcessorAlgos/histogram_median_algo.h compiled unmodified) on a seeded input.  The reference
ships no golden vectors of its own, so these are the pins: the C restatement
(oracle/median_oracle.c) and the CUDA path must reproduce them bit for bit on any box.

    python tests/golden/make_golden.py
"""
import ctypes
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from cvvidproc_b200 import synth  # noqa: E402

SIG = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]


def ref_median(frames, bin_bytes=0):
    lib = ctypes.CDLL(str(REPO / "oracle" / "_ref" / "libcvvp_median_ref.so"))
    fn = lib.cvvp_ref_median
    fn.argtypes = SIG
    fn.restype = ctypes.c_int
    frames = np.ascontiguousarray(frames)
    nelem = int(np.prod(frames.shape[1:]))
    out = np.empty(nelem, np.uint8)
    assert fn(frames.ctypes.data, frames.shape[0], nelem, nelem, bin_bytes, 4, out.ctypes.data) == 0
    return out.reshape(frames.shape[1:])


def case_input(case):
    kind = case["kind"]
    if kind == "random":
        rng = np.random.default_rng(case["seed"])
        return rng.integers(case["lo"], case["hi"], (case["n"], case["h"], case["w"]), dtype=np.uint8)
    if kind == "synth":
        return synth.synth_frames(0, case["n"], case["w"], case["h"], case["seed"], case["ndisks"])
    if kind == "two_valued":
        st = np.empty((case["n"], case["h"], case["w"]), np.uint8)
        st[: case["n_low"]] = case["low"]
        st[case["n_low"]:] = case["high"]
        return st
    if kind == "constant":
        return np.full((case["n"], case["h"], case["w"]), case["value"], np.uint8)
    raise ValueError(kind)


CASES = [
    dict(name="rand_n1", kind="random", seed=1, n=1, h=9, w=13, lo=0, hi=256),
    dict(name="rand_n2", kind="random", seed=2, n=2, h=9, w=13, lo=0, hi=256),
    dict(name="rand_n3", kind="random", seed=3, n=3, h=9, w=13, lo=0, hi=256),
    dict(name="rand_n100", kind="random", seed=100, n=100, h=31, w=47, lo=0, hi=256),
    dict(name="rand_n101", kind="random", seed=101, n=101, h=31, w=47, lo=0, hi=256),
    dict(name="rand_n255_u8bins", kind="random", seed=255, n=255, h=16, w=33, lo=0, hi=256),
    dict(name="rand_n256_u16bins", kind="random", seed=256, n=256, h=16, w=33, lo=0, hi=256),
    dict(name="rand_n1000_narrow", kind="random", seed=1000, n=1000, h=8, w=40, lo=100, hi=124),
    dict(name="rand_ragged_641x479", kind="random", seed=7, n=25, h=479, w=641, lo=0, hi=256),
    dict(name="two_valued_even_split", kind="two_valued", n=100, n_low=50, low=10, high=250, h=4, w=8),
    dict(name="two_valued_low_majority", kind="two_valued", n=100, n_low=51, low=10, high=250, h=4, w=8),
    dict(name="constant_200", kind="constant", n=77, value=200, h=5, w=5),
    # saturation of the reference's u8 counters: 300 identical frames forced through HistogramMedianAlgo8
    dict(name="saturated_u8_constant", kind="constant", n=300, value=42, h=3, w=3, bin_bytes=1),
    dict(name="saturated_u8_two_valued", kind="two_valued", n=600, n_low=290, low=7, high=9, h=3, w=3, bin_bytes=1),
    dict(name="synth_c1_640x480x100", kind="synth", seed=1, ndisks=4, n=100, h=480, w=640),
    dict(name="synth_c2_reduced_480x270x300", kind="synth", seed=2, ndisks=30, n=300, h=270, w=480),
]


def main():
    out = []
    for case in CASES:
        frames = case_input(case)
        res = ref_median(frames, case.get("bin_bytes", 0))
        entry = dict(case)
        entry["input_sha256"] = hashlib.sha256(frames.tobytes()).hexdigest()
        entry["output_sha256"] = hashlib.sha256(res.tobytes()).hexdigest()
        if res.size <= 64:
            entry["output"] = res.reshape(-1).tolist()
        out.append(entry)
        print(case["name"], entry["output_sha256"][:16])
    (Path(__file__).parent / "median_golden.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
