"""Writes profiles/sass_median_pipe.txt: which SASS instructions prove that the median kernels are TMA + mbarrier
pipelines (B200_PROFILING.md "What proves a Blackwell-native kernel"): cp.async.bulk.tensor shows up as UTMALDG, mbarrier
try_wait / arrive as SYNCS.PHASECHK.TRANS64.TRYWAIT / SYNCS.ARRIVE.TRANS64*.  Runs on the CPU box (cuobjdump only).
    python tools/sass_evidence.py
"""
import collections
import re
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
PATTERNS = ["UTMALDG", r"LDSM\.8\.MT1616", r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT", r"SYNCS\.ARRIVE\.TRANS64", "POPC", "LOP3", "PRMT", "ATOMS", "MEMBAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(REPO / "cvvidproc_b200" / "libcvvp_cuda.so")], check=True, capture_output=True,
                         text=True).stdout
    rows, tot = [], collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        m = re.search(r"median_pipe_kernelILi(\d)ELi(\d)ELi(\d+)ELi(\d)E", name)
        if not m:
            continue
        key = f"median_pipe_kernel<LOG2S={m.group(1)},JT={m.group(2)},NSELW={m.group(3)},MODE={m.group(4)}>"
        c = {k: len(re.findall(k, f)) for k in PATTERNS}
        rows.append((key, c))
        tot.update(c)
    clean = lambda k: k.replace("\\", "")  # noqa: E731
    out = ["# SASS evidence for the median kernels (cuobjdump -sass cvvidproc_b200/libcvvp_cuda.so, sm_100a; tools/sass_evidence.py)",
           "# TMA (cp.async.bulk.tensor) appears as UTMALDG, mbarrier try_wait / arrive as SYNCS.PHASECHK.TRANS64.TRYWAIT /",
           "# SYNCS.ARRIVE.TRANS64*.  No UTC*MMA / LDTM is expected: nothing on this path is a contraction.",
           f"# {len(rows)} instantiations of median_pipe_kernel; totals: " + ", ".join(f"{clean(k)}={v}" for k, v in tot.items()), "",
           "# the variants the 1080p x 1000-frame jobs run (LOG2S=0, JT=8, NSELW=16): MODE 0 = on-chip select, 1 / 2 = nibble",
           "# counting rounds, 3 = window counting"]
    for key, c in rows:
        if "LOG2S=0,JT=8,NSELW=16" in key:
            out.append(f"{key}: " + ", ".join(f"{clean(k)}={v}" for k, v in c.items()))
    out += ["", "# every instantiation"]
    for key, c in sorted(rows):
        out.append(f"{key}: " + " ".join(f"{clean(k).split('.')[-1]}={v}" for k, v in c.items()))
    (REPO / "profiles" / "sass_median_pipe.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out[:12]))


if __name__ == "__main__":
    main()
