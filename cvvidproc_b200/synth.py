"""Deterministic synthetic uint8 frames -- host twin of csrc/synth.cu.

Integer-only arithmetic, so these bytes are identical to what cvvp_synth_frames_device writes.
Used by the parity tests (small sizes) and to build host-resident inputs.

    mix32    = murmur3 fmix32
    r(f,y,x) = mix32(seed ^ mix32(f*0x9E3779B1 + (y*W + x)))
    B(y,x)   = 140 + (x*20)//W - (y*10)//H ;  noise = (r & 7) - 3
    disk k   : a = mix32(seed*1000003 + k), b = mix32(a), c = mix32(b)
               cx = (a % W + 2f) % W, cy = b % H, rad = 3 + c % 30, depth = 10 + (c>>8) % 50
               core (rad > 8): radius rad//3, adds back 10 + (c>>16) % 50
    frame    = clamp(B + noise - sum(depth in disks) + sum(core add-back), 0, 255)
"""
from __future__ import annotations

import numpy as np

# seeds / disk counts of the five BASELINE.json configs
CONFIG_PARAMS = {
    "C1": dict(width=640, height=480, nframes=100, seed=1, ndisks=4),
    "C2": dict(width=1920, height=1080, nframes=1000, seed=2, ndisks=30),
    "C3": dict(width=1920, height=1080, nframes=10000, seed=3, ndisks=30),
    "C4": dict(width=512, height=256, nframes=200000, seed=4, ndisks=6),
    "C5": dict(width=3840, height=2160, nframes=5000, seed=5, ndisks=60),
}


def mix32(h):
    h = np.asarray(h, dtype=np.uint32).copy()
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def _mix32_scalar(h: int) -> int:
    h &= 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def disk_params(seed: int, k: int, f: int, width: int, height: int):
    a = _mix32_scalar(seed * 1000003 + k)
    b = _mix32_scalar(a)
    c = _mix32_scalar(b)
    cx = (a % width + 2 * f) % width
    cy = b % height
    rad = 3 + c % 30
    depth = 10 + (c >> 8) % 50
    if rad > 8:
        core_r = rad // 3
        core_add = 10 + (c >> 16) % 50
    else:
        core_r = -1
        core_add = 0
    return cx, cy, rad, depth, core_r, core_add


def synth_frame(f: int, width: int, height: int, seed: int, ndisks: int, row0: int = 0, nrows: int | None = None) -> np.ndarray:
    if nrows is None:
        nrows = height - row0
    with np.errstate(over="ignore"):
        ys = np.arange(row0, row0 + nrows, dtype=np.int64)[:, None]
        xs = np.arange(width, dtype=np.int64)[None, :]
        lin = (ys * width + xs).astype(np.uint32)
        fterm = np.uint32((f * 0x9E3779B1) & 0xFFFFFFFF)
        r = mix32(np.uint32(seed & 0xFFFFFFFF) ^ mix32(lin + fterm))
    v = 140 + (xs * 20) // width - (ys * 10) // height + (r & np.uint32(7)).astype(np.int64) - 3
    for k in range(ndisks):
        cx, cy, rad, depth, core_r, core_add = disk_params(seed, k, f, width, height)
        y_lo, y_hi = max(cy - rad, row0), min(cy + rad, row0 + nrows - 1)
        if y_lo > y_hi:
            continue
        x_lo, x_hi = max(cx - rad, 0), min(cx + rad, width - 1)
        yy = np.arange(y_lo, y_hi + 1)[:, None]
        xx = np.arange(x_lo, x_hi + 1)[None, :]
        q = (xx - cx) ** 2 + (yy - cy) ** 2
        sub = v[y_lo - row0 : y_hi - row0 + 1, x_lo : x_hi + 1]
        sub -= np.where(q <= rad * rad, depth, 0)
        if core_r >= 0:
            sub += np.where(q <= core_r * core_r, core_add, 0)
    return np.clip(v, 0, 255).astype(np.uint8)


def synth_frames(first_frame: int, nframes: int, width: int, height: int, seed: int, ndisks: int, row0: int = 0,
                 nrows: int | None = None) -> np.ndarray:
    if nrows is None:
        nrows = height - row0
    out = np.empty((nframes, nrows, width), np.uint8)
    for i in range(nframes):
        out[i] = synth_frame(first_frame + i, width, height, seed, ndisks, row0, nrows)
    return out


# The only highlight parameter set the reference ever uses (Sources/rand_tests.cpp:42-51, :333-342): the benchmark's
# workload definition.  The structuring element is cv2.getStructuringElement(MORPH_ELLIPSE, (4, 4)) as cv2 4.13 produces
# it, spelled out so that nothing depends on an OpenCV version's ellipse rasteriser.
CANONICAL_HIGHLIGHT = dict(
    struct_element=((0, 0, 1, 0), (1, 1, 1, 1), (1, 1, 1, 1), (1, 1, 1, 1)),
    threshold=14, threshold_lo=7, threshold_hi=16, min_size_hyst=20, min_size_threshold=20, width_border=5,
)


def canonical_struct_element() -> np.ndarray:
    return np.array(CANONICAL_HIGHLIGHT["struct_element"], np.uint8)

