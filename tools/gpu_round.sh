#!/bin/bash
# One GPU-box pass of what the driver runs at round end: the GPU parity suite, smoke, both bench arms.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','parity_spot_check')}, d['roofline']['frac'], d['e2e']['value'])
for k in ('highlight','frame_source','c5_median','c4_highlight','track_e2e'):
    print(k, d[k].get('value'), d[k].get('parity_spot_check'))
"
