"""Developer probe (GPU box, CPU only): how well does cv2 decode in one Python thread overlap with a cv2 tracker callback in
another?  Prints frames/s of each alone and together."""
import sys, time, threading, tempfile
from pathlib import Path
import numpy as np
import cv2

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from cvvidproc_b200 import synth
import ctypes

W, H, N = 1920, 1080, 300
lib = ctypes.CDLL(str(REPO / "oracle" / "_build" / "libcvvp_oracle.so"))
fn = lib.cvvp_oracle_synth_frames
fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_int] * 4 + [ctypes.c_longlong] * 2 + [ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
frames = np.empty((N, H, W), np.uint8)
fn(frames.ctypes.data, H * W, W, H, 0, H, 0, N, 3, 30, 16)
d = tempfile.mkdtemp()
path = str(Path(d) / "v.avi")
vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (W, H), isColor=True)
for f in frames:
    vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
vw.release()
mask = ((frames[0] < 120) * 255).astype(np.uint8)
bufs = np.empty((32, H, W, 3), np.uint8)

def decode(out):
    cap = cv2.VideoCapture(path)
    t0 = time.perf_counter(); k = 0
    while cap.read(bufs[k % 32])[0]:
        k += 1
    out["dec"] = (time.perf_counter() - t0) / k * 1e3

def ccl(out, n):
    t0 = time.perf_counter()
    for i in range(n):
        nn, _, stats, cent = cv2.connectedComponentsWithStats(mask, connectivity=8)
        _ = [(int(stats[j, cv2.CC_STAT_AREA]), round(float(cent[j][0]), 2)) for j in range(1, nn)]
    out["ccl"] = (time.perf_counter() - t0) / n * 1e3

for nthreads in (None, 1):
    if nthreads is not None:
        cv2.setNumThreads(nthreads)
    o = {}
    decode(o); ccl(o, 100)
    print(f"cv2 threads {cv2.getNumThreads()}: alone: decode {o['dec']:.2f} ms/frame (32 rotating buffers), ccl {o['ccl']:.2f} ms/call", flush=True)
    o = {}
    ta = threading.Thread(target=decode, args=(o,)); tb = threading.Thread(target=ccl, args=(o, 150))
    t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join()
    print(f"   together: decode {o['dec']:.2f} ms/frame, ccl {o['ccl']:.2f} ms/call, wall {time.perf_counter()-t0:.2f} s", flush=True)
