#!/bin/bash
# ncu --set full capture of one fused-highlight launch (296 frames of 1080p) with source counters
CMD="python tools/prof_highlight.py C3 296"
$CMD > gpurun_out/plain_hl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:highlight_fused -s 3 -c 1 -o gpurun_out/prof_hl_fused $CMD > gpurun_out/ncu_hl.log 2>&1
tail -3 gpurun_out/ncu_hl.log
