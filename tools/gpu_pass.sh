#!/bin/bash
# generic GPU pass: bash tools/gpu_pass.sh "<pytest args>" [probe commands ...]
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest $1 -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
shift
i=0
for cmd in "$@"; do
  i=$((i+1))
  echo "== $cmd"
  timeout 600 bash -c "$cmd" > gpurun_out/probe_$i.log 2>&1; tail -25 gpurun_out/probe_$i.log
done
