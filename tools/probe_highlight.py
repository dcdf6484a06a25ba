"""Device-resident timing of the highlight stage (development probe)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from cvvidproc_b200 import _cabi, synth
from oracle import highlight_oracle as ho


def main():
    ctx = _cabi.Context(0)
    cases = [("C3", 1920, 1080, 64, 3, 30), ("C4", 512, 256, 2048, 4, 6)]
    for name, w, h, n, seed, k in cases:
        npix = w * h
        frames = torch.empty((n, npix), dtype=torch.uint8, device="cuda:0")
        out = torch.empty((n, npix), dtype=torch.uint8, device="cuda:0")
        ctx.synth_frames_device(frames.data_ptr(), npix, w, h, 0, n, seed, k)
        bgstack = torch.empty((255, npix), dtype=torch.uint8, device="cuda:0")
        bg = torch.empty(npix, dtype=torch.uint8, device="cuda:0")
        ctx.synth_frames_device(bgstack.data_ptr(), npix, w, h, 0, 255, seed, k)
        ctx.median_device(bgstack.data_ptr(), 255, npix, npix, bg.data_ptr())
        ctx.synchronize()
        bg_h = bg.cpu().numpy().reshape(h, w)
        p = ho.canonical_params(bg_h)
        ctx.highlight_begin(p.background, p.struct_element, p.threshold, p.threshold_lo, p.threshold_hi, p.min_size_hyst,
                            p.min_size_threshold, p.width_border)
        stream = torch.cuda.ExternalStream(ctx.stream)
        with torch.cuda.stream(stream):
            for _ in range(2):
                ctx.highlight_device(frames.data_ptr(), n, npix, out.data_ptr(), npix)
            l0 = ctx.launch_count
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            reps = 5
            for _ in range(reps):
                ctx.highlight_device(frames.data_ptr(), n, npix, out.data_ptr(), npix)
            b.record(stream)
        ctx.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{name} {w}x{h} x{n}: {ms:.3f} ms per batch = {ms / n * 1e3:.1f} us/frame -> "
              f"{n * npix / 1e6 / ms * 1e3:.0f} Mpx-frames/s, {(ctx.launch_count - l0) // reps} launches/batch", flush=True)
        # parity spot check on a few frames
        got = out.cpu().numpy().reshape(n, h, w)
        fr = frames.cpu().numpy().reshape(n, h, w)
        bad = 0
        for i in (0, n // 2, n - 1):
            bad += int((ho.highlight_objects(fr[i].copy(), p) != got[i]).sum())
        print("   spot-check:", "OK" if bad == 0 else f"MISMATCH {bad}", "white frac", float((got == 255).mean()), flush=True)
        ctx.highlight_end()
        del frames, out, bgstack
    ctx.close()


if __name__ == "__main__":
    main()
