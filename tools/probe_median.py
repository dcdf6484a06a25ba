"""Quick device-resident timing of the median kernel (development probe, not the benchmark)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from cvvidproc_b200 import _cabi


def main():
    ctx = _cabi.Context(0)
    cases = [(1920, 1080, 1000), (1920, 1080, 1024), (1920, 1080, 512), (1920, 1080, 625), (640, 480, 100),
             (1920, 1080, 2000), (3840, 2160, 625)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for (w, h, n) in cases:
        nelem = w * h
        stack = torch.empty((n, nelem), dtype=torch.uint8, device="cuda:0")
        out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
        ctx.synth_frames_device(stack.data_ptr(), nelem, w, h, 0, n, 2, 30)
        ctx.synchronize()
        stream = torch.cuda.ExternalStream(ctx.stream)
        with torch.cuda.stream(stream):
            for _ in range(3):
                ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in evs:
                a.record(stream)
                ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
                b.record(stream)
        ctx.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        gb = n * nelem / 1e9
        print(f"{w}x{h}x{n}: median {med:.3f} ms  min {ms[0]:.3f} ms  -> {gb / med * 1e3:.0f} GB/s  "
              f"({gb / ms[0] * 1e3:.0f} best), {n * nelem / 1e6 / med * 1e3:.3e} Mpx-frames/s", flush=True)
        # spot check
        cols = np.arange(0, nelem, max(1, nelem // 2000))
        samp = stack[:, torch.from_numpy(cols).cuda()].cpu().numpy()
        ok = np.array_equal(np.sort(samp, axis=0)[n // 2], out.cpu().numpy()[cols])
        print("   spot-check:", "OK" if ok else "MISMATCH", flush=True)
        del stack, out
    ctx.close()


if __name__ == "__main__":
    main()
