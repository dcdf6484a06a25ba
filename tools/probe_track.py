"""End-to-end probe of the Python surface on a lossless synthetic video: GetVideoBackground + TrackObjects with a no-op
tracker, (a) frame preparation on the host and one batch in flight (the synchronous shape), (b) device preparation and
the asynchronous queue (the default).  Decode (cv2, FFV1) is inside the timed region here -- this is the user's view.
    python tools/probe_track.py [nframes] [C4|C3]
"""
import os
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import numpy as np  # noqa: E402

import cvvidproc_b200 as cvp  # noqa: E402
import video_util  # noqa: E402
from cvvidproc_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 480
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
    p = synth.CONFIG_PARAMS[cfg]
    w, h = p["width"], p["height"]
    frames = synth.synth_frames(0, n, w, h, p["seed"], p["ndisks"])
    with tempfile.TemporaryDirectory() as d:
        path = video_util.write_lossless(Path(d) / "c4.avi", frames)
        t0 = time.perf_counter()
        import cv2

        cap = cv2.VideoCapture(path)
        k = 0
        while cap.read()[0]:
            k += 1
        t_dec = time.perf_counter() - t0
        print(f"decode alone: {k} frames of {w}x{h} in {t_dec * 1e3:.0f} ms ({t_dec / k * 1e3:.2f} ms/frame)", flush=True)
        calls = [0]

        def tracker(bw_frame, frames_processed, objects_prev, objects_archive, next_ID, kwargs):
            calls[0] += 1
            objects_archive[frames_processed] = int(bw_frame[::8, ::8].any())
            return next_ID

        cp = synth.CANONICAL_HIGHLIGHT
        os.environ["CVVP_TRACK_BATCH"] = "32" if cfg == "C4" else "8"

        def run(env, limit, nbg, ntrack):
            os.environ.pop("CVVP_HOST_PREP", None)
            os.environ.update(env)
            t0 = time.perf_counter()
            bg = cvp.GetVideoBackground(cvp.VidBgPack(path, vid_is_grayscale=True, frame_limit=nbg))
            t1 = time.perf_counter()
            hp = cvp.HighlightObjectsPack(bg, synth.canonical_struct_element(), cp["threshold"], cp["threshold_lo"],
                                          cp["threshold_hi"], cp["min_size_hyst"], cp["min_size_threshold"], cp["width_border"])
            calls[0] = 0
            arch = cvp.TrackObjects(cvp.VidObjectTrackPack(path, hp, cvp.AssignObjectsPack(tracker, {}), vid_is_grayscale=True,
                                                           token_storage_limit=limit, frame_limit=ntrack))
            t2 = time.perf_counter()
            assert calls[0] == ntrack and len(arch) == ntrack
            return t1 - t0, t2 - t1

        run({}, 10, 8, 8)  # warm-up: CUDA context, module load
        for label, env, limit in (("host prep, 1 batch in flight ", {"CVVP_HOST_PREP": "1"}, 1),
                                  ("host prep, queue depth 3     ", {"CVVP_HOST_PREP": "1"}, 10),
                                  ("device prep, queue depth 3   ", {}, 10)):
            tb, tt = run(env, limit, min(n, 255), n)
            m = n
            print(f"{label}: background {1e3 * tb:.0f} ms, TrackObjects {1e3 * tt:.0f} ms "
                  f"({tt / m * 1e3:.2f} ms/frame, {m * w * h / 1e6 / tt:.0f} Mpx-frames/s incl. decode and callback)", flush=True)


if __name__ == "__main__":
    main()
