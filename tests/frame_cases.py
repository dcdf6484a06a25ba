"""Seeded decoded-frame cases for the frame source (crop + channel reduction): shared by the oracle tests, the golden
generator and the GPU parity tests.  Each case: (name, frames (n, H, W[, C]) uint8, crop (x, y, w, h), mode)."""
import numpy as np

AS_IS, CHANNEL0, RGB2GRAY = 0, 1, 2


def _rand(seed, shape):
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def cases():
    out = []
    # whole frames, every mode, 3 channels (what cv::VideoCapture hands out)
    for mode, tag in ((AS_IS, "asis"), (CHANNEL0, "ch0"), (RGB2GRAY, "gray")):
        out.append((f"full_64x48x3_{tag}", _rand(1, (3, 48, 64, 3)), (0, 0, 64, 48), mode))
    # crop offsets that leave every source alignment (x*3 mod 16) and odd output widths
    for k, (x, y, w, h) in enumerate([(1, 0, 61, 47), (5, 3, 33, 20), (7, 11, 1, 1), (0, 47, 64, 1), (63, 0, 1, 48),
                                      (2, 2, 17, 9), (13, 1, 50, 46)]):
        for mode, tag in ((CHANNEL0, "ch0"), (RGB2GRAY, "gray"), (AS_IS, "asis")):
            out.append((f"crop{k}_{tag}", _rand(10 + k, (2, 48, 64, 3)), (x, y, w, h), mode))
    # rows wider than one 1024-element tile, with a ragged tail, cropped and not
    out.append(("wide_2500_gray", _rand(30, (2, 5, 2500, 3)), (0, 0, 2500, 5), RGB2GRAY))
    out.append(("wide_2500_crop_gray", _rand(31, (2, 5, 2500, 3)), (3, 1, 2049, 3), RGB2GRAY))
    out.append(("wide_2500_asis", _rand(32, (1, 4, 2500, 3)), (1, 0, 2400, 4), AS_IS))
    # 4 channels (alpha ignored), 1 channel (already grey), 2 channels as is
    out.append(("rgba_gray", _rand(40, (2, 30, 41, 4)), (2, 3, 37, 20), RGB2GRAY))
    out.append(("rgba_ch0", _rand(41, (2, 30, 41, 4)), (0, 0, 41, 30), CHANNEL0))
    out.append(("mono_ch0", _rand(42, (3, 30, 41)), (4, 5, 30, 21), CHANNEL0))
    out.append(("mono_asis", _rand(43, (3, 30, 41)), (0, 0, 41, 30), AS_IS))
    out.append(("two_asis", _rand(44, (2, 19, 23, 2)), (1, 1, 20, 17), AS_IS))
    # saturated / extreme colours: rounding at both ends of the fixed-point range
    ext = np.zeros((1, 4, 8, 3), np.uint8)
    ext[0, 0] = 255
    ext[0, 1, :, 0] = 255
    ext[0, 2, :, 1] = 255
    ext[0, 3, :, 2] = 255
    out.append(("extremes_gray", ext, (0, 0, 8, 4), RGB2GRAY))
    # reduced 1080p-like geometry
    out.append(("hd_480x270_gray", _rand(50, (4, 270, 480, 3)), (0, 0, 480, 270), RGB2GRAY))
    out.append(("hd_480x270_crop_ch0", _rand(51, (4, 270, 480, 3)), (40, 30, 400, 200), CHANNEL0))
    return out
