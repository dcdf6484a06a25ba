#!/bin/bash
set -u
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
tail -4 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().splitlines()[-1])
print(len(open('gpurun_out/bench_n1.json').read().splitlines()), 'stdout lines')
print(json.dumps(d['track_e2e'], indent=1)); print(d['highlight']['e2e'], d['c5_median']['ms_per_step'])
"
P3="python tools/probe_median.py 1920x1080x5000"
$P3 > gpurun_out/plain_p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:median_pipe -s 30 -c 1 -o gpurun_out/r2_prof_median_mode3 $P3 > gpurun_out/ncu_p3.log 2>&1; echo "mode3 rc=$?"
