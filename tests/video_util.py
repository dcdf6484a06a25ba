"""Lossless test videos (FFV1 in AVI round-trips bit-exactly through cv2 in this image)."""
import cv2
import numpy as np


def write_lossless(path, frames, fps=30.0):
    """frames: (n, H, W) gray or (n, H, W, 3) BGR uint8."""
    frames = np.asarray(frames)
    color = frames.ndim == 4
    h, w = frames.shape[1:3]
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), fps, (w, h), isColor=True)
    if not vw.isOpened():
        raise RuntimeError("cannot open FFV1 writer")
    for f in frames:
        vw.write(f if color else cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    return str(path)


def read_all(path):
    cap = cv2.VideoCapture(str(path))
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    return np.stack(out)
