"""Times the frame-source kernel (csrc/frames.cu) alone on device-resident decoded frames: geometry / mode sweep.
    python tools/probe_frames.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cvvidproc_b200 import _cabi  # noqa: E402


def run(ctx, stream, w, h, c, mode, crop, n, reps=10):
    src = torch.randint(0, 256, (n, h * w * c), dtype=torch.uint8, device="cuda:0")
    fmt = _cabi.FrameFormat.of((h, w, c) if c > 1 else (h, w), mode, crop)
    ob = int(torch.tensor(fmt.out_shape).prod())
    dst = torch.empty((n, ob), dtype=torch.uint8, device="cuda:0")
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.frames_prepare_device(src.data_ptr(), n, h * w * c, fmt, dst.data_ptr(), ob)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            ctx.frames_prepare_device(src.data_ptr(), n, h * w * c, fmt, dst.data_ptr(), ob)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    cw, ch = (crop[2], crop[3]) if crop else (w, h)
    step = c if mode != _cabi.FRAMES_AS_IS else 1
    elems = cw * ch * (c if mode == _cabi.FRAMES_AS_IS else 1)
    gb = n * elems * (step + 1) / 1e9
    print(f"{w}x{h}x{c} mode {mode} crop {crop}: {ms:.4f} ms / {n} frames, {n * cw * ch / 1e6 / (ms * 1e-3):.0f} Mpx-frames/s, "
          f"{gb / (ms * 1e-3):.0f} GB/s algorithmic")


def main():
    ctx = _cabi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    G, C0, A = _cabi.FRAMES_RGB2GRAY, _cabi.FRAMES_CHANNEL0, _cabi.FRAMES_AS_IS
    run(ctx, stream, 1920, 1080, 3, G, None, 96)
    run(ctx, stream, 1920, 1080, 3, C0, None, 96)
    run(ctx, stream, 1920, 1080, 3, A, None, 64)
    run(ctx, stream, 1920, 1080, 3, G, (17, 9, 1801, 1000), 96)
    run(ctx, stream, 3840, 2160, 3, G, None, 24)
    run(ctx, stream, 512, 256, 3, G, None, 1500)
    run(ctx, stream, 1920, 1080, 1, C0, (100, 100, 1600, 900), 200)
    ctx.close()


if __name__ == "__main__":
    main()
