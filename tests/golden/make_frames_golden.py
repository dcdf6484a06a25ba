"""Generates tests/golden/frames_golden.json: SHA-256 of the frame-source oracle's output (oracle/frames_oracle.py,
cv2 4.13.0) on every case of tests/frame_cases.py.  The reference has no fixtures for this path; these hashes pin the
oracle (and through it the CUDA path) against drift.

    python tests/golden/make_frames_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import frame_cases  # noqa: E402
from oracle import frames_oracle as fo  # noqa: E402


def main():
    out = []
    for name, frames, crop, mode in frame_cases.cases():
        res = fo.prepare_frames(frames, crop, mode)
        out.append(dict(name=name, shape=list(frames.shape), crop=list(crop), mode=mode, out_shape=list(res.shape),
                        input_sha256=hashlib.sha256(frames.tobytes()).hexdigest(),
                        output_sha256=hashlib.sha256(res.tobytes()).hexdigest()))
    (Path(__file__).parent / "frames_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print(len(out), "entries")


if __name__ == "__main__":
    main()
