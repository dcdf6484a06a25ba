#!/bin/bash
# bench.py under torchrun on N GPUs of one box:  bash tools/gpu_bench_n.sh N [extra bench args]
set -u
N=$1; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
tail -8 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.json').read().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'parity_spot_check', 'undecided_elements') if k in d})
    print(d['roofline'].get('phase_ms'), d['e2e']['value'])
    for k in ('highlight', 'c5_median', 'c4_highlight'):
        if k in d:
            print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['e2e']['value'], d[k].get('parity_spot_check'), d[k].get('roofline', {}).get('phase_ms'))
except Exception as e:
    print('no line:', e)
PY
