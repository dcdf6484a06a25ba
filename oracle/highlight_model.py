"""TEST INFRASTRUCTURE ONLY -- label-based model of the highlight stage (the executable spec of the CUDA kernels).

Same outputs as oracle/highlight_oracle.py (the cv2 restatement of highlight_objects_algo.cpp), but expressed
with connected-component labels and purely local rules instead of contour tracing, polygon areas, polygon filling
and flood fills -- i.e. in the form the GPU kernels compute.  numpy + scipy.ndimage only.
tests/test_oracle_highlight.py holds this model to the cv2 restatement on random and adversarial images; the CUDA
kernels are then held to both.

Conventions: S8 / S4 = 8- / 4-connectivity; "FRAME" = the background region connected to the (zero-padded) image
exterior; raster order = row-major.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

S8 = np.ones((3, 3), bool)
S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)


def diff_sat(background: np.ndarray, frame: np.ndarray) -> np.ndarray:
    """9.2: d = max(int(bg) - int(frame), 0)  (highlight_objects_algo.cpp:27-29)"""
    return np.clip(background.astype(np.int16) - frame.astype(np.int16), 0, 255).astype(np.uint8)


def otsu_threshold(d: np.ndarray) -> int:
    """9.6: OpenCV's getThreshVal_Otsu_8u restated (double arithmetic, first maximum wins)."""
    h = np.bincount(d.reshape(-1), minlength=256).astype(np.float64)
    n = float(d.size)
    scale = 1.0 / n
    mu = float(np.dot(np.arange(256, dtype=np.float64), h)) * scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def threshold_mask(d: np.ndarray, t: int) -> np.ndarray:
    """thr(d,t) = [d > t]; t == -1 -> Otsu (ThresholdImage :81-104)"""
    if t == -1:
        t = otsu_threshold(d)
    return d.astype(np.int32) > int(t)


def morph_open(mask: np.ndarray, selem: np.ndarray) -> np.ndarray:
    """9.2: erode then dilate with the SAME (unreflected) offsets, anchor (kw//2, kh//2); out-of-image samples are
    ignored (erode of nothing = set, dilate of nothing = clear).  cv::morphologyEx(MORPH_OPEN) :39, :61"""
    kh, kw = selem.shape
    ay, ax = kh // 2, kw // 2
    offs = [(i - ay, j - ax) for i in range(kh) for j in range(kw) if selem[i, j] != 0]
    if not offs:
        # OpenCV quirk: a kernel without non-zero entries is filtered as if only its element (0,0) were set
        # (preprocess2DKernel sizes the coordinate list to max(nz,1) and leaves its single entry at the origin)
        offs = [(-ay, -ax)]
    h, w = mask.shape

    def shifted(img, dy, dx, fill):
        out = np.full_like(img, fill)
        ys0, ys1 = max(0, -dy), min(h, h - dy)
        xs0, xs1 = max(0, -dx), min(w, w - dx)
        if ys0 < ys1 and xs0 < xs1:
            out[ys0:ys1, xs0:xs1] = img[ys0 + dy : ys1 + dy, xs0 + dx : xs1 + dx]
        return out

    er = np.ones_like(mask)
    for dy, dx in offs:
        er &= shifted(mask, dy, dx, True)
    di = np.zeros_like(mask)
    for dy, dx in offs:
        di |= shifted(er, dy, dx, False)
    return di


def _first_pixels(labels: np.ndarray, nlab: int):
    """raster-first flat index of every label 1..nlab"""
    flat = labels.reshape(-1)
    first = np.full(nlab + 1, -1, np.int64)
    idx = np.flatnonzero(flat)
    # reversed assignment: the smallest index written last wins
    first[flat[idx[::-1]]] = idx[::-1]
    return first


def hysteresis(d: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """9.3 (ThresholdImageWithHysteresis :107-144)"""
    di = d.astype(np.int32)
    U = di > hi
    L = di > lo
    h, w = d.shape
    Up = np.pad(U, 1)
    fgU, nfg = ndimage.label(Up, structure=S8)
    bgU, _ = ndimage.label(~Up, structure=S4)
    frame_lab = bgU[0, 0]
    L1, _ = ndimage.label(L, structure=S4)
    L0, _ = ndimage.label(~L, structure=S4)
    out = np.zeros((h, w), bool)
    if nfg == 0:
        return out
    first = _first_pixels(fgU, nfg)
    wp = w + 2
    keep1 = set()
    keep0 = set()
    for c in range(1, nfg + 1):
        y, x = divmod(int(first[c]), wp)
        if bgU[y, x - 1] != frame_lab:
            continue  # RETR_EXTERNAL: components inside a hole of another component are not seeds
        sy, sx = y - 1, x - 1  # contour[0] = raster-first pixel, unpadded
        if L[sy, sx]:
            keep1.add(int(L1[sy, sx]))
        else:  # only possible when lo > hi: the seed sits on a zero pixel of the lower mask
            keep0.add(int(L0[sy, sx]))
    if keep1:
        out |= np.isin(L1, list(keep1))
    if keep0:
        out |= np.isin(L0, list(keep0))
    return out


def remove_small_objects(mask: np.ndarray, min_size: int) -> np.ndarray:
    """9.4 (RemoveSmallObjects :146-181): contour polygon areas from crack sums and convex-corner counts, the
    drawContours edge rule, and the even-odd nesting parity of the single filled draw."""
    h, w = mask.shape
    M = np.pad(mask.astype(bool), 1)
    hp, wp = M.shape
    fg, nfg = ndimage.label(M, structure=S8)
    bg, nbg = ndimage.label(~M, structure=S4)
    if nfg == 0:
        return mask.copy()
    frame_lab = int(bg[0, 0])
    key_mul = nbg + 1

    s = np.zeros((nfg + 1) * key_mul, np.int64)
    E = np.zeros_like(s)
    Xv = np.zeros_like(s)
    xs = np.arange(wp)[None, :].repeat(hp, 0)

    def acc(arr, sel_fg, sel_bg, weights):
        keys = fg[sel_fg].astype(np.int64) * key_mul + bg[sel_bg]
        np.add.at(arr, keys, weights)

    # left / right cracks
    m = M[:, 1:] & ~M[:, :-1]  # fg at x, bg at x-1
    yy, xx = np.nonzero(m)
    acc(s, (yy, xx + 1), (yy, xx), -(xs[yy, xx + 1]))
    acc(E, (yy, xx + 1), (yy, xx), 1)
    m = M[:, :-1] & ~M[:, 1:]  # fg at x, bg at x+1
    yy, xx = np.nonzero(m)
    acc(s, (yy, xx), (yy, xx + 1), xs[yy, xx] + 1)
    acc(E, (yy, xx), (yy, xx + 1), 1)
    # up / down cracks
    m = M[1:, :] & ~M[:-1, :]
    yy, xx = np.nonzero(m)
    acc(E, (yy + 1, xx), (yy, xx), 1)
    m = M[:-1, :] & ~M[1:, :]
    yy, xx = np.nonzero(m)
    acc(E, (yy, xx), (yy + 1, xx), 1)
    # convex corners: 2x2 blocks with exactly one fg pixel
    a, b_, c, d_ = M[:-1, :-1], M[:-1, 1:], M[1:, :-1], M[1:, 1:]
    cnt = a.astype(np.int8) + b_ + c + d_
    one = cnt == 1
    for sel, (fy, fx), (by, bx) in (
        (one & a, (0, 0), (0, 1)),
        (one & b_, (0, 1), (0, 0)),
        (one & c, (1, 0), (0, 0)),
        (one & d_, (1, 1), (0, 0)),
    ):
        yy, xx = np.nonzero(sel)
        acc(Xv, (yy + fy, xx + fx), (yy + by, xx + bx), 1)

    Lc = E - Xv
    twoA = np.where(s > 0, 2 * s - Lc - 2, 2 * np.abs(s) + Lc - 2)
    small = (E > 0) & (twoA < 2 * int(min_size))

    # helpers on the nesting tree
    first_fg = _first_pixels(fg, nfg)
    first_bg = _first_pixels(bg, nbg)
    b_out = np.zeros(nfg + 1, np.int64)
    for cc in range(1, nfg + 1):
        y, x = divmod(int(first_fg[cc]), wp)
        b_out[cc] = bg[y, x - 1]
    parent = np.zeros(nbg + 1, np.int64)
    for bb in range(1, nbg + 1):
        if bb == frame_lab:
            continue
        y, x = divmod(int(first_bg[bb]), wp)
        parent[bb] = fg[y, x - 1]

    def is_small(cc, bb):
        return bool(small[cc * key_mul + bb])

    depth_odd = np.zeros(nfg + 1, bool)
    for cc in range(1, nfg + 1):
        dcount = 0
        cur = cc
        while True:
            bo = int(b_out[cur])
            if not is_small(cur, bo):
                break
            dcount += 1
            if bo == frame_lab:
                break
            par = int(parent[bo])
            if not is_small(par, bo):
                break
            dcount += 1
            cur = par
        depth_odd[cc] = (dcount & 1) == 1

    # zeroing: edge rule, then parity rule
    zero = np.zeros_like(M)
    for dy, dx in ((0, -1), (0, 1), (-1, 0), (1, 0)):
        ys0, ys1 = max(0, -dy), hp - max(0, dy)
        xs0, xs1 = max(0, -dx), wp - max(0, dx)
        f = M[ys0:ys1, xs0:xs1] & ~M[ys0 + dy : ys1 + dy, xs0 + dx : xs1 + dx]
        keys = fg[ys0:ys1, xs0:xs1].astype(np.int64) * key_mul + bg[ys0 + dy : ys1 + dy, xs0 + dx : xs1 + dx]
        zero[ys0:ys1, xs0:xs1] |= f & small[np.where(f, keys, 0)]
    zero |= M & depth_odd[fg]
    out = M & ~zero
    return out[1:-1, 1:-1]


def fill_holes(mask: np.ndarray) -> np.ndarray:
    """9.5 (FillHoles :183-221)"""
    h, w = mask.shape
    m = mask.astype(bool)
    seed = (0, 0) if m[0, 0] else (h - 1, w - 1)
    if m[seed]:
        return np.ones_like(m)  # flood fill from a set pixel is a no-op -> NOT(clone) | image = all set
    bgl, _ = ndimage.label(~m, structure=S4)
    return ~(bgl == bgl[seed])


def highlight_objects(frame: np.ndarray, background: np.ndarray, struct_element: np.ndarray, threshold: int,
                      threshold_lo: int, threshold_hi: int, min_size_hyst: int, min_size_threshold: int,
                      stages: dict | None = None) -> np.ndarray:
    d = diff_sat(background, frame)
    selem = np.asarray(struct_element) != 0
    a0 = threshold_mask(d, threshold)
    a1 = morph_open(a0, selem)
    a2 = remove_small_objects(a1, min_size_threshold)
    a3 = fill_holes(a2)
    b0 = hysteresis(d, threshold_lo, threshold_hi)
    b1 = morph_open(b0, selem)
    b2 = remove_small_objects(b1, min_size_hyst)
    b3 = fill_holes(b2)
    if stages is not None:
        to8 = lambda m: (m.astype(np.uint8) * 255)
        stages.update(diff=d, a_thresh=to8(a0), a_open=to8(a1), a_rso=to8(a2), a_fill=to8(a3), b_hyst=to8(b0),
                      b_open=to8(b1), b_rso=to8(b2), b_fill=to8(b3))
    return ((a3 | b3).astype(np.uint8)) * 255
