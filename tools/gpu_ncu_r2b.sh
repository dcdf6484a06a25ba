#!/bin/bash
# round 2, final evidence pass: launch list of the bench command + --set full of the fused highlight kernel (1024 frames of 1080p)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-track"
if $CMD > gpurun_out/plain_bench.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
fi
HL="python tools/prof_highlight.py C3 1024"
if $HL > gpurun_out/plain_hl.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:highlight_fused -s 3 -c 1 -o gpurun_out/r2_prof_hl_fused $HL > gpurun_out/ncu_hl.log 2>&1; echo "highlight rc=$?"
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
