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
