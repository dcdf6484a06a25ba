#!/bin/bash
# round 2, pass A: the window-counting median (parity + phase times)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_median_shard_gpu.py -m gpu -x -q > gpurun_out/pytest_shard.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_shard.log
for w in 1 2 8; do timeout 300 python tools/probe_shard.py $w >> gpurun_out/probe_shard.log 2>&1; done
timeout 300 python tools/probe_median.py 1920x1080x1000 1920x1080x5000 3840x2160x2500 > gpurun_out/probe_median.log 2>&1; cat gpurun_out/probe_median.log
tail -5 gpurun_out/pytest_shard.log; cat gpurun_out/probe_shard.log
