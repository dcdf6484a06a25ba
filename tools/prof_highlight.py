"""Developer aid (GPU box): time the fused highlight kernel on device-resident synthetic frames and print its per-phase
profile (CVVP_HL_PROF=1).   python tools/prof_highlight.py [C3|C4] [nframes] [slots ...]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from cvvidproc_b200 import _cabi, synth  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    slot_list = [int(v) for v in sys.argv[3:]] or [0]
    p_ = synth.CONFIG_PARAMS[cfg]
    W, H = p_["width"], p_["height"]
    npix = W * H
    for slots in slot_list:
        if slots:
            os.environ["CVVP_HL_SLOTS"] = str(slots)
        else:
            os.environ.pop("CVVP_HL_SLOTS", None)
        with _cabi.Context(0) as ctx:
            stream = torch.cuda.ExternalStream(ctx.stream)
            bgstack = torch.empty((255, npix), dtype=torch.uint8, device="cuda")
            bg = torch.empty(npix, dtype=torch.uint8, device="cuda")
            ctx.synth_frames_device(bgstack.data_ptr(), npix, W, H, 0, 255, p_["seed"], p_["ndisks"])
            ctx.median_device(bgstack.data_ptr(), 255, npix, npix, bg.data_ptr())
            ctx.synchronize()
            del bgstack
            cp = synth.CANONICAL_HIGHLIGHT
            frames = torch.empty((n, npix), dtype=torch.uint8, device="cuda")
            masks = torch.empty((n, npix), dtype=torch.uint8, device="cuda")
            ctx.synth_frames_device(frames.data_ptr(), npix, W, H, 1000, n, p_["seed"], p_["ndisks"])
            ctx.highlight_begin(bg.cpu().numpy().reshape(H, W), synth.canonical_struct_element(), cp["threshold"],
                                cp["threshold_lo"], cp["threshold_hi"], cp["min_size_hyst"], cp["min_size_threshold"],
                                cp["width_border"])
            os.environ.pop("CVVP_HL_PROF", None)
            for _ in range(2):
                ctx.highlight_device(frames.data_ptr(), n, npix, masks.data_ptr(), npix)
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(3):
                    ctx.highlight_device(frames.data_ptr(), n, npix, masks.data_ptr(), npix)
                e1.record(stream)
            ctx.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"{cfg} {W}x{H} n={n} in_flight={ctx.highlight_frames_in_flight()}: {ms:.3f} ms per batch, "
                  f"{ms * 1e3 / n:.2f} us/frame, {npix * n / 1e6 / (ms * 1e-3):.0f} Mpx-frames/s, "
                  f"{2 * npix * n / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)
            os.environ["CVVP_HL_PROF"] = "1"
            ctx.highlight_device(frames.data_ptr(), n, npix, masks.data_ptr(), npix)
            ctx.synchronize()
            os.environ.pop("CVVP_HL_PROF", None)
            ctx.highlight_end()


if __name__ == "__main__":
    main()
