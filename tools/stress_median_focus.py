"""Development aid: repeated runs of the on-chip select in chosen variants; prints where it disagrees with torch.sort
(tile iteration of the CTA, element inside the tile).
    python tools/stress_median_focus.py reps"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from cvvidproc_b200 import _cabi

CASES = [  # (label, n, env, P)
    ("A log2s=1 half JT5", 1248, {}, 64),
    ("B log2s=1 half JT3 forced", 600, {"CVVP_MEDIAN_LOG2S": "1", "CVVP_MEDIAN_BUFFERS": "1"}, 64),
    ("C log2s=1 two buffers", 1000, {"CVVP_MEDIAN_LOG2S": "1"}, 64),
    ("D log2s=1 half JT8", 2000, {"CVVP_MEDIAN_TWO_PASS": "0"}, 64),
    ("E log2s=0 half JT5", 530, {}, 128),
    ("E log2s=0 half JT7", 769, {}, 128),
    ("E log2s=0 half JT8", 1000, {}, 128),
    ("F log2s=0 two buffers", 300, {}, 128),
    ("G log2s=2 half", 2500, {"CVVP_MEDIAN_TWO_PASS": "0"}, 32),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    ctx = _cabi.Context(0)
    for label, n, env, P in CASES:
        for k in ("CVVP_MEDIAN_LOG2S", "CVVP_MEDIAN_BUFFERS", "CVVP_MEDIAN_TWO_PASS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for nelem in (128 * 600, 128 * 2000):
            nbad = 0
            its, els, diffs = [], [], []
            for rep in range(reps):
                g = torch.Generator(device="cuda:0").manual_seed(n * 131 + nelem * 7 + rep)
                stack = torch.randint(0, 256, (n, nelem), dtype=torch.uint8, device="cuda:0", generator=g)
                out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
                torch.cuda.synchronize()  # the library launches on its own stream: the generator must be done
                ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
                ctx.synchronize()
                want = torch.sort(stack, dim=0).values[n // 2]
                bad = (out != want).nonzero().flatten().cpu().numpy()
                if bad.size:
                    nbad += 1
                    tiles = bad // P
                    its += (tiles // 148).tolist()
                    els += (bad % P).tolist()
                    diffs += (out[bad].int() - want[bad].int()).cpu().numpy().tolist()
                del stack, out, want
            msg = f"{label}: n={n} nelem={nelem} ({nelem // P} tiles of {P}): {nbad}/{reps} runs wrong"
            if nbad:
                msg += (f"; {len(its)} elements; by tile iteration {np.bincount(its).tolist()[:16]}; "
                        f"by element {np.bincount(els, minlength=P).tolist()}; diffs {np.unique(diffs).tolist()}")
            print(msg, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
