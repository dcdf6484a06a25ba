"""Generates tests/golden/highlight_golden.json: SHA-256 of the REFERENCE's output -- HighlightObjectsAlgo compiled
unmodified from /root/reference against the cv2-forwarding shim (oracle/_ref/cvvp_highlight_ref, OpenCV 4.13.0 through
the cv2 wheel) -- on every adversarial case, 40 random cases and reduced synthetic-stream frames.  The reference has no
fixtures of its own for this path; these are outputs of the reference itself run here.  The cv2 restatement
(oracle/highlight_oracle.py) must produce the same bytes, or the script stops.  Run where /root/reference is mounted:

    python tests/golden/make_highlight_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import hl_cases  # noqa: E402
from oracle import highlight_oracle as ho  # noqa: E402
from oracle import highlight_ref as href  # noqa: E402


def entry(kind, name, frame, p, **extra):
    out = href.highlight_objects(frame, p)
    if not (out == ho.highlight_objects(frame.copy(), p)).all():
        raise SystemExit(f"{name}: the restatement differs from the compiled reference")
    e = dict(kind=kind, name=name, shape=list(frame.shape), input_sha256=hashlib.sha256(frame.tobytes()).hexdigest(),
             output_sha256=hashlib.sha256(out.tobytes()).hexdigest(), white_fraction=float((out == 255).mean()),
             source="reference (oracle/_ref/cvvp_highlight_ref)")
    e.update(extra)
    return e


def main():
    out = []
    for name, frame, p in hl_cases.adversarial_cases():
        out.append(entry("adversarial", name, frame, p))
    for t in range(40):
        frame, p = hl_cases.random_case(t)
        out.append(entry("random", f"random_{t}", frame, p, t=t))
    for cfg, fi, scale in (("C3", 7, 4), ("C3", 200, 4), ("C4", 3, 1), ("C4", 5000, 1)):
        frame, p = hl_cases.synthetic_case(cfg, fi, scale)
        out.append(entry("synthetic", f"synth_{cfg}_{fi}_s{scale}", frame, p, cfg=cfg, frame_index=fi, scale=scale))
    (Path(__file__).parent / "highlight_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print(len(out), "entries")


if __name__ == "__main__":
    main()
