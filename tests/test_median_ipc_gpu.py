"""The frame-sharded median across PROCESSES: two ranks, one process each, exchange buffers mapped through
cvvp_median_shard_export / cvvp_median_shard_import (cudaIpcOpenMemHandle), barriers through torch.distributed.
This is the path `bench.py --gpus N` takes under torchrun; the other shard tests attach emulated ranks inside one process.
With two GPUs every rank gets its own device (real peer stores over NVLink); with one, both ranks share it -- no kernel
of the library waits for another rank, so that is legal (B200_PROFILING.md) and still crosses the process boundary."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CVVP_REPO"])
from cvvidproc_b200 import _cabi, sharded

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
dev = rank % ngpu
torch.cuda.set_device(dev)
dist.init_process_group("gloo")  # handles and barriers only; the data path is peer memory
n, nelem = 700, 5000
rng = np.random.default_rng(11)
frames = rng.integers(118, 126, (n, nelem), dtype=np.uint8)
frames[:, :700] = rng.integers(0, 256, (n, 700), dtype=np.uint8)  # undecidable in one pass: the two-round exchange runs too
first, cnt = sharded.frame_chunk(n, rank, world)
ctx = _cabi.Context(dev)
stride = (nelem + 127) // 128 * 128
stack = torch.zeros((cnt, stride), dtype=torch.uint8, device=f"cuda:{dev}")
stack[:, :nelem] = torch.from_numpy(frames[first:first + cnt]).to(f"cuda:{dev}")
torch.cuda.synchronize()

def barrier():  # host barrier: this rank's phase is complete, then wait for the others
    ctx.synchronize()
    dist.barrier()

job = sharded.ShardedMedian(ctx, nelem, rank, world, barrier=barrier, max_rank_frames=cnt)
job.connect_processes()
want = np.sort(frames, axis=0)[n // 2]
out = {}
# easy part only: decided in one pass
left = job.run_window(stack.data_ptr(), cnt, stride)
got = ctx.copy_to_host(job.result_ptr(), nelem)
out["window_left"] = int(left)
out["window_ok"] = bool(np.array_equal(got[700:], want[700:]))
job.run(stack.data_ptr(), cnt, stride)
got = ctx.copy_to_host(job.result_ptr(), nelem)
out["full_ok"] = bool(np.array_equal(got, want))
job.run_two_round(stack.data_ptr(), cnt, stride)
out["two_round_ok"] = bool(np.array_equal(ctx.copy_to_host(job.result_ptr(), nelem), want))
dist.barrier()
job.close()
ctx.close()
print("RESULT", rank, out, flush=True)
dist.destroy_process_group()
'''


def test_two_processes_exchange_through_ipc_handles(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CVVP_REPO=str(REPO), MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        line = [ln for ln in o.splitlines() if ln.startswith("RESULT")][-1]
        res = eval(line.split(" ", 2)[2])
        assert res["window_ok"] and res["full_ok"] and res["two_round_ok"], (rank, res)
        assert 0 < res["window_left"] <= 700, (rank, res)
