"""Developer probe: TrackObjects on a lossless 1080p video with the per-thread time breakdown (CVVP_TRACK_DEBUG=1)."""
import os, sys, time, tempfile
from pathlib import Path
import numpy as np
import cv2
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import cvvidproc_b200 as cvp
from cvvidproc_b200 import synth
import ctypes
W, H, N = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 400
lib = ctypes.CDLL(str(REPO / "oracle" / "_build" / "libcvvp_oracle.so"))
fn = lib.cvvp_oracle_synth_frames
fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_int] * 4 + [ctypes.c_longlong] * 2 + [ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
frames = np.empty((N, H, W), np.uint8)
fn(frames.ctypes.data, H * W, W, H, 0, H, 0, N, 3, 30, 16)
d = tempfile.mkdtemp(); path = str(Path(d) / "v.avi")
vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (W, H), isColor=True)
for f in frames: vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
vw.release()
bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True, frame_limit=255))
cp = synth.CANONICAL_HIGHLIGHT
hp = cvp.HighlightObjectsPack(bg, synth.canonical_struct_element(), cp["threshold"], cp["threshold_lo"], cp["threshold_hi"], cp["min_size_hyst"], cp["min_size_threshold"], cp["width_border"])
def noop(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs): return next_ID
def ccl(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
    n, _, stats, cent = cv2.connectedComponentsWithStats(bw_frame, connectivity=8)
    objects_archive[frames_processed] = [(int(stats[i, cv2.CC_STAT_AREA]), round(float(cent[i][0]), 2)) for i in range(1, n)]
    return next_ID + n - 1
os.environ["CVVP_TRACK_DEBUG"] = "1"
for name, fn_ in (("noop", noop), ("noop", noop), ("ccl", ccl), ("ccl", ccl)):
    for batch in ("", "4", "64"):
        if batch: os.environ["CVVP_TRACK_BATCH"] = batch
        else: os.environ.pop("CVVP_TRACK_BATCH", None)
        t0 = time.perf_counter()
        cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(fn_, {}), vid_is_grayscale=True))
        dt = time.perf_counter() - t0
        print(f"{name} batch={batch or 'default'}: {dt/N*1e3:.2f} ms/frame", flush=True)
