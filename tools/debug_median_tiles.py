"""Development aid: where does the on-chip select disagree with numpy?  (frame counts x element counts)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from cvvidproc_b200 import _cabi


def main():
    ctx = _cabi.Context(0)
    ns = [int(a) for a in sys.argv[1:]] or [769, 780, 800, 896, 897]
    for n in ns:
        for nelem in (128 * 148, 128 * 149, 128 * 296, 128 * 469):
            g = torch.Generator(device="cuda:0").manual_seed(n)
            stack = torch.randint(60, 200, (n, nelem), dtype=torch.uint8, device="cuda:0", generator=g)
            out = torch.empty(nelem, dtype=torch.uint8, device="cuda:0")
            torch.cuda.synchronize()  # the library launches on its own stream: the generator must be done
            ctx.median_device(stack.data_ptr(), n, nelem, nelem, out.data_ptr())
            ctx.synchronize()
            want = torch.sort(stack, dim=0).values[n // 2]
            bad = (out != want).nonzero().flatten().cpu().numpy()
            tiles = np.unique(bad // 128)
            print(f"n={n} nelem={nelem} ({nelem // 128} tiles): {bad.size} wrong elements in {tiles.size} tiles", end="")
            if bad.size:
                rounds = np.unique(tiles // 148)
                print(f"; tile rounds {rounds[:10]}; first tiles {tiles[:8]}; elements in tile {np.unique(bad % 128)[:16]}")
            else:
                print()
    ctx.close()


if __name__ == "__main__":
    main()
