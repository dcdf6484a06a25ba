"""Development probe: median kernel bandwidth against the frame stride (same 1920x1080 element range, 1000 frames)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cvvidproc_b200 import _cabi


def main():
    ctx = _cabi.Context(0)
    n, nelem = 1000, 1920 * 1080
    strides = [2073600, 2073600 + 128, 2097152, 2097152 + 4096, 2621440, 3145728, 3686400, 4147200, 4194304, 4194304 + 65536,
               6220800, 8294400, 8388608, 8388608 + 2097152 // 2]
    if len(sys.argv) > 1:
        strides = [int(a) for a in sys.argv[1:]]
    buf = torch.zeros(max(strides) * n, dtype=torch.uint8, device="cuda:0")
    buf.random_(0, 256)
    out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.ExternalStream(ctx.stream)
    for s in strides:
        with torch.cuda.stream(stream):
            for _ in range(3):
                ctx.median_device(buf.data_ptr(), n, nelem, s, out.data_ptr())
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in evs:
                a.record(stream)
                ctx.median_device(buf.data_ptr(), n, nelem, s, out.data_ptr())
                b.record(stream)
        ctx.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        print(f"stride {s:9d} ({s / 2097152:.3f} x 2 MiB): {med:.3f} ms -> {n * nelem / 1e6 / med:.0f} GB/s", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
