#!/bin/bash
# round 2 ncu evidence (one GPU; every ncu pass only after the same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-track"
if $CMD > gpurun_out/plain_bench.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
P0="python tools/probe_median.py 1920x1080x1000"
if $P0 > gpurun_out/plain_p0.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:median_pipe -s 4 -c 1 -o gpurun_out/r2_prof_median_mode0 $P0 > gpurun_out/ncu_p0.log 2>&1; echo "mode0 rc=$?"
fi
P3="python tools/probe_median.py 1920x1080x5000"
if $P3 > gpurun_out/plain_p3.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:median_pipe -s 22 -c 1 -o gpurun_out/r2_prof_median_mode3 $P3 > gpurun_out/ncu_p3.log 2>&1; echo "mode3 rc=$?"
  ncu --set full --clock-control none -k regex:shard_window_final -s 4 -c 1 -o gpurun_out/r2_prof_window_final $P3 > gpurun_out/ncu_p3b.log 2>&1; echo "final rc=$?"
fi
ls -la gpurun_out | tail -12
