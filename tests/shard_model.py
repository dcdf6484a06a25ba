"""numpy model of the four phases of csrc/median_shard.cu (test infrastructure: it lets the world_size-2 gloo tests
walk the exchange protocol -- who owns which elements, what is pushed where, what the owner decides -- on CPU).
The product never imports this."""
import numpy as np


def nibble_counts(frames: np.ndarray, nibble: str, sel: np.ndarray | None = None) -> np.ndarray:
    """frames (n, nelem) uint8 -> (nelem, 16) uint16 counts of the high nibble, or of the low nibble among the frames
    whose high nibble equals sel[e] & 15 (phases 0 and 2)."""
    n, nelem = frames.shape
    out = np.zeros((nelem, 16), np.uint16)
    if n == 0:
        return out
    hi, lo = frames >> 4, frames & 15
    for b in range(16):
        if nibble == "hi":
            out[:, b] = (hi == b).sum(axis=0)
        else:
            out[:, b] = ((lo == b) & (hi == (sel & 15)[None, :])).sum(axis=0)
    return out


def owner_pick_hi(counts: np.ndarray) -> np.ndarray:
    """counts (world, owned, 16) -> sel (owned,) uint32 = h | k' << 8 (phase 1)."""
    c = counts.astype(np.int64).sum(axis=0)
    total = c.sum(axis=1)
    k = total // 2
    cum = np.cumsum(c, axis=1)
    h = (cum > k[:, None]).argmax(axis=1)
    below = np.where(h > 0, cum[np.arange(len(h)), np.maximum(h - 1, 0)], 0)
    return (h | ((k - below) << 8)).astype(np.uint32)


def owner_pick_lo(counts: np.ndarray, sel: np.ndarray) -> np.ndarray:
    """counts (world, owned, 16), sel (owned,) -> result bytes (phase 3)."""
    c = counts.astype(np.int64).sum(axis=0)
    k = (sel >> 8).astype(np.int64)
    cum = np.cumsum(c, axis=1)
    low = (cum > k[:, None]).argmax(axis=1)
    return (((sel & 15) << 4) | low).astype(np.uint8)


# ---- the one-pass form (phases 4 and 5 of csrc/median_shard.cu) ------------------------------------------------------
def pilot_median(frames: np.ndarray) -> np.ndarray:
    """what a source (one launch of <= 1024 frames) centres its window on.  The kernel takes the upper median of two of its
    plane rows (256 frames); ANY pilot gives the same final image -- it only decides which elements are settled in one
    pass -- so the model uses the upper median of frames [0, 128) + [n/2, n/2 + 128), the frames those rows hold."""
    n = frames.shape[0]
    pick = np.r_[0:min(n, 128), min(n, n // 2):min(n, n // 2 + 128)]
    sub = frames[np.unique(pick)]
    return np.sort(sub, axis=0)[sub.shape[0] // 2]


def window_records(frames: np.ndarray) -> np.ndarray:
    """frames (n, nelem) uint8 of ONE source -> (nelem, 11) int64 records: 8 bins of the window [base, base + 7], frames
    below the window, the window base, the source's frame count (phase 4; the kernel packs them into 32 bytes)."""
    n, nelem = frames.shape
    rec = np.zeros((nelem, 11), np.int64)
    if n == 0:
        return rec
    c = pilot_median(frames).astype(np.int64)
    base = np.where(c >= 4, np.minimum(c - 4, 248), 0)
    d = frames.astype(np.int64) - base[None, :]
    for b in range(8):
        rec[:, b] = (d == b).sum(axis=0)
    rec[:, 8] = (d < 0).sum(axis=0)
    rec[:, 9] = base
    rec[:, 10] = n
    return rec


def owner_window_final(records: np.ndarray):
    """records (sources, owned, 11) -> (result bytes (owned,), decided (owned,) bool) (phase 5): the cumulative count of all
    sources is exact on the intersection [lo - 1, hi] of their windows; the median is the first value there whose
    cumulative count exceeds N / 2 (histogram_median_algo.h:160-166) -- if it lies there at all."""
    nsrc, owned, _ = records.shape
    res = np.zeros(owned, np.uint8)
    ok = np.zeros(owned, bool)
    for e in range(owned):
        live = [records[s, e] for s in range(nsrc) if records[s, e, 10] > 0]
        if not live:
            continue
        total = sum(int(r[10]) for r in live)
        lo = max(int(r[9]) for r in live)
        hi = min(int(r[9]) + 7 for r in live)
        if lo > hi:
            continue
        k = total // 2

        def cum(v):  # frames of all sources with value <= v, exact for lo - 1 <= v <= hi
            return sum(int(r[8]) + int(r[: max(0, v - int(r[9]) + 1)].sum()) for r in live)

        if cum(lo - 1) > k:
            continue
        for v in range(lo, hi + 1):
            if cum(v) > k:
                res[e], ok[e] = v, True
                break
    return res, ok
