#!/bin/bash
# one ncu --set full capture: bash tools/gpu_ncu_one.sh <kernel regex> <skip> <output name> <command ...>
# (the command first runs plain; ncu only after it exited 0)
set -u
mkdir -p gpurun_out
K=$1; S=$2; O=$3; shift 3
if "$@" > gpurun_out/plain_$O.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/$O "$@" > gpurun_out/ncu_$O.log 2>&1; echo "ncu rc=$?"
else
  echo "plain run failed"; tail -20 gpurun_out/plain_$O.log
fi
tail -5 gpurun_out/plain_$O.log
ls -la gpurun_out/$O.ncu-rep
