"""Find the first intermediate stage at which the fused highlight kernel leaves the cv2 oracle (GPU box only).

    python tools/debug_highlight_stages.py            # runs the shared test cases, prints the first bad stage of each
The kernel's CVVP_HL_DEBUG_STAGE hook writes the bit image of stage k instead of the final mask.
"""
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import hl_cases  # noqa: E402
from cvvidproc_b200 import _cabi  # noqa: E402
from oracle import highlight_oracle as ho  # noqa: E402

STAGES = {1: "a_thresh", 4: "a_open", 5: "a_rso", 6: "a_fill", 7: "b_hyst", 8: "b_open", 9: "b_rso", 10: "b_fill"}


def run(ctx, frame, p, stage):
    if stage:
        os.environ["CVVP_HL_DEBUG_STAGE"] = str(stage)
    else:
        os.environ.pop("CVVP_HL_DEBUG_STAGE", None)
    ctx.highlight_begin(p.background, np.ascontiguousarray(p.struct_element), p.threshold, p.threshold_lo, p.threshold_hi,
                        p.min_size_hyst, p.min_size_threshold, p.width_border)
    try:
        return ctx.highlight_frames(frame[None])[0]
    finally:
        ctx.highlight_end()
        os.environ.pop("CVVP_HL_DEBUG_STAGE", None)


def main():
    cases = [(f"random{t}", *hl_cases.random_case(t)) for t in range(60)] + hl_cases.adversarial_cases()
    for h, w in [(40, 32), (33, 64), (30, 160), (17, 640)]:
        f, bg = hl_cases.blob_frame(h, w, 300, sigma=2.0, amp=70)
        cases.append((f"aligned{h}x{w}", f, ho.canonical_params(bg)))
    nbad = 0
    with _cabi.Context(0) as ctx:
        for name, frame, p in cases:
            stages = {}
            want = ho.highlight_objects(frame.copy(), p, stages)
            got = run(ctx, frame, p, 0)
            if np.array_equal(got, want):
                continue
            nbad += 1
            msg = f"{name} {frame.shape}: final differs in {(got != want).sum()} px;"
            for k, key in STAGES.items():
                g = run(ctx, frame, p, k)
                if not np.array_equal(g, stages[key]):
                    ys, xs = np.nonzero(g != stages[key])
                    msg += f" first bad stage {k} ({key}): {len(ys)} px, e.g. (y={ys[0]}, x={xs[0]}) got {g[ys[0], xs[0]]}"
                    break
            else:
                msg += " every stage matches (final OR / white flags?)"
            print(msg)
    print(f"{nbad} of {len(cases)} cases differ")


if __name__ == "__main__":
    main()
