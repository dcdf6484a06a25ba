"""Generates tests/golden/frames_golden.json: SHA-256 of the frame source's output on every case of tests/frame_cases.py.
Every case with 3-channel frames (what cv::VideoCapture hands out) is written to a lossless FFV1 video and run through
the REFERENCE's own generator -- CvVidFramesGeneratorAlgo compiled unmodified from /root/reference against the
cv2-forwarding shim (oracle/_ref/cvvp_frames_ref, OpenCV 4.13.0 through the cv2 wheel): those hashes are outputs of the
reference itself (source = "reference"), and the cv2 restatement (oracle/frames_oracle.py) must produce the same bytes
or the script stops.  Cases a decoder never produces (1, 2, 4 channels) and the two
5-row cases the video writer truncates are hashed from the restatement (source = "restatement ...").  Run where /root/reference is mounted:

    python tests/golden/make_frames_golden.py
"""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import frame_cases  # noqa: E402
import video_util  # noqa: E402
from oracle import frames_oracle as fo  # noqa: E402
from oracle import frames_ref as fref  # noqa: E402


def main():
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, frames, crop, mode in frame_cases.cases():
            res = fo.prepare_frames(frames, crop, mode)
            source = "restatement"
            if frames.ndim == 4 and frames.shape[3] == 3:
                vid = video_util.write_lossless(Path(tmp) / f"{name}.avi", frames)
                decoded = video_util.read_all(vid)
                if decoded.shape == frames.shape and np.array_equal(decoded, frames):
                    toks = fref.tokens(vid, 0, len(frames), crop, mode, frames_in_batch=3)
                    if len(toks) != len(frames) or not np.array_equal(np.stack(toks), res):
                        raise SystemExit(f"{name}: the restatement differs from the compiled reference")
                    source = "reference (oracle/_ref/cvvp_frames_ref)"
                else:  # an odd frame height loses its last row in cv2's FFV1 writer
                    source = "restatement (the geometry does not survive the video writer)"
            out.append(dict(name=name, shape=list(frames.shape), crop=list(crop), mode=mode, out_shape=list(res.shape),
                            input_sha256=hashlib.sha256(frames.tobytes()).hexdigest(),
                            output_sha256=hashlib.sha256(res.tobytes()).hexdigest(), source=source))
    (Path(__file__).parent / "frames_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print(len(out), "entries,", sum(e["source"].startswith("reference") for e in out), "from the reference")


if __name__ == "__main__":
    main()
