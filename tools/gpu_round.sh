#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), probes, ncu launch list and --set full captures of the
# dominant kernels (each ncu pass only after the same command has exited 0 without ncu).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/probe_frames.py > gpurun_out/probe_frames.log 2>&1
python tools/prof_highlight.py C3 1024 > gpurun_out/prof_hl_c3.log 2>&1
python tools/prof_highlight.py C4 16384 > gpurun_out/prof_hl_c4.log 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
if $CMD > gpurun_out/plain.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:median_pipe -s 3 -c 1 -o gpurun_out/prof_median $CMD > gpurun_out/ncu_median.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:highlight_fused -s 3 -c 1 -o gpurun_out/prof_highlight $CMD > gpurun_out/ncu_highlight.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:frames_prepare -s 3 -c 1 -o gpurun_out/prof_frames $CMD > gpurun_out/ncu_frames.log 2>&1
fi
ls -la gpurun_out
