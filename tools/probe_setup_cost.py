"""Developer probe: fixed costs of one TrackObjects-shaped job at 1080p (context, parameters, queue, scratch, teardown)."""
import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from cvvidproc_b200 import _cabi, synth
W, H = 1920, 1080
bg = np.full((H, W), 140, np.uint8)
cp = synth.CANONICAL_HIGHLIGHT
fr = np.full((17, H, W, 3), 130, np.uint8)
def t(label, fn):
    t0 = time.perf_counter(); r = fn(); print(f"{label:28s} {1e3 * (time.perf_counter() - t0):8.1f} ms", flush=True); return r
for rep in range(2):
    print("--- pass", rep)
    ctx = t("ctx create", lambda: _cabi.Context(0))
    t("highlight_begin", lambda: ctx.highlight_begin(bg, synth.canonical_struct_element(), cp["threshold"], cp["threshold_lo"], cp["threshold_hi"], cp["min_size_hyst"], cp["min_size_threshold"], cp["width_border"]))
    fmt = _cabi.FrameFormat.of((H, W, 3), _cabi.FRAMES_CHANNEL0)
    t("queue_begin depth 3 x 17", lambda: ctx.highlight_queue_begin(3, 17, fmt))
    t("first submit (scratch alloc)", lambda: (ctx.highlight_submit(fr), ctx.highlight_next()))
    t("second submit", lambda: (ctx.highlight_submit(fr), ctx.highlight_next()))
    t("queue_end", ctx.highlight_queue_end)
    t("highlight_end", ctx.highlight_end)
    t("ctx close", ctx.close)
