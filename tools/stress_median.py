"""Repeated-run sweep of the median kernels over adversarial geometry (development aid; compute-sanitizer's racecheck
is closed on this pool, so races in the mbarrier / counter protocol have to show as wrong medians):
every stage count of the on-chip select (1 .. 40 stages of 32 frames, +-1 frame around each boundary), the counting
path beyond it, tile counts below / at / above the SM count so that CTAs walk 0, 1 and several tiles, three runs each
with different data, every result against torch.sort on the device.

    python tools/stress_median.py [repeats]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from cvvidproc_b200 import _cabi


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ctx = _cabi.Context(0)
    ns = sorted({max(1, 32 * k + d) for k in range(1, 41) for d in (-1, 0, 1)} | {1, 2, 3, 5, 17, 100, 1000})
    ns += [1281, 1500, 2047, 2048, 2049, 2100, 3000, 4100]
    nelems = [128 * 3 + 16, 128 * 147, 128 * 148, 128 * 149 + 64, 128 * 469]
    t0 = time.time()
    runs = bad_runs = 0
    for n in ns:
        for nelem in nelems:
            if n > 1280 and nelem > 128 * 149 + 64 and n > 2100:
                continue  # keep the long stacks small: the sort dominates
            for rep in range(reps):
                g = torch.Generator(device="cuda:0").manual_seed(n * 131 + nelem * 7 + rep)
                lo, hi = ((60, 200), (0, 256), (120, 136))[rep % 3]
                stack = torch.randint(lo, hi, (n, nelem), dtype=torch.uint8, device="cuda:0", generator=g)
                out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
                torch.cuda.synchronize()  # the library launches on its own stream: the generator must be done
                ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
                ctx.synchronize()
                want = torch.sort(stack, dim=0).values[n // 2]
                bad = (out != want).nonzero().flatten()
                runs += 1
                if bad.numel():
                    bad_runs += 1
                    b = bad.cpu().numpy()
                    print(f"MISMATCH n={n} nelem={nelem} rep={rep}: {b.size} elements, tiles {np.unique(b // 128)[:8]}", flush=True)
                del stack, out, want
    ctx.close()
    print(f"{runs} runs over {len(ns)} frame counts x {len(nelems)} element counts x {reps} repeats: {bad_runs} with wrong medians "
          f"({time.time() - t0:.1f} s)")
    sys.exit(1 if bad_runs else 0)


if __name__ == "__main__":
    main()
