"""Development probe: does the median kernel lose bandwidth over long launches?  Times one launch over the whole image
against the same image cut into K element ranges launched back to back on one stream."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cvvidproc_b200 import _cabi


def main():
    ctx = _cabi.Context(0)
    cases = [(3840, 2160, 1000), (1920, 1080, 1000)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for (w, h, n) in cases:
        nelem = w * h
        stack = torch.empty((n, nelem), dtype=torch.uint8, device="cuda:0")
        out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
        ctx.synth_frames_device(stack.data_ptr(), nelem, w, h, 0, n, 2, 30)
        ctx.synchronize()
        stream = torch.cuda.ExternalStream(ctx.stream)
        for pieces in (1, 2, 4, 8, 16):
            per = (nelem // pieces + 127) // 128 * 128
            def run():
                for i in range(pieces):
                    first = i * per
                    cnt = min(per, nelem - first)
                    if cnt > 0:
                        ctx.median_device(stack.data_ptr() + first, n, cnt, nelem, out.data_ptr() + first)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    run()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a, b in evs:
                    a.record(stream)
                    run()
                    b.record(stream)
            ctx.synchronize()
            ms = sorted(a.elapsed_time(b) for a, b in evs)
            med = ms[len(ms) // 2]
            print(f"{w}x{h}x{n} in {pieces:2d} launches: {med:.3f} ms -> {n * nelem / 1e6 / med:.0f} GB/s", flush=True)
        del stack, out
    ctx.close()


if __name__ == "__main__":
    main()
