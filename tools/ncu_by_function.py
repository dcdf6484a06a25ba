"""Aggregate an ncu SASS source page by source function (CPU-side helper).

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libcvvp_cuda.so; nvdisasm -g -c <cubin> > fused.sass
    python tools/ncu_by_function.py sass.csv fused.sass cvvidproc_b200/csrc/highlight_fused.cu
Maps every SASS instruction to the source line nvdisasm reports (-lineinfo build), then sums executed warp instructions
and stall samples per enclosing function of the .cu file.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, disasm, cu = sys.argv[1:4]
    per_line = len(sys.argv) > 4
    # function start lines
    starts = []
    for i, line in enumerate(open(cu), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__)[^(]*?\b(\w+)\s*\(", line)
        if m:
            starts.append((i, m.group(1)))
        elif re.match(r"^\s+(?:__device__)[^(]*?\b(\w+)\s*\(", line):  # methods
            starts.append((i, re.match(r"^\s+(?:__device__)[^(]*?\b(\w+)\s*\(", line).group(1)))

    def func_of(line):
        name = "?"
        for s, n in starts:
            if s <= line:
                name = n
            else:
                break
        return name

    # offset -> line
    off2line = {}
    cur = None
    in_kernel = False
    for l in open(disasm):
        if ".text." in l and "highlight_fused_kernel" in l:
            in_kernel = True
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith(cu.split("/")[-1]) else -1
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m and in_kernel:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr = rows[h]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = defaultdict(lambda: [0, 0])
    for r in rows[h + 1:]:
        try:
            addr = int(r[ia], 16)
            n = int(r[ii])
            sm = int(r[isamp])
        except Exception:
            continue
        if base is None:
            base = addr
        line = off2line.get(addr - base)
        key = (line if per_line else func_of(line)) if line and line > 0 else "other"
        agg[key][0] += n
        agg[key][1] += sm
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[1] for v in agg.values())
    print(f"total warp instructions {tot}, samples {tots}")
    for k, (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"{100 * n / tot:6.2f}% instr  {100 * sm / max(tots, 1):6.2f}% samples  {k}")


if __name__ == "__main__":
    main()
