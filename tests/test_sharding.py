"""Host-side multi-rank logic on CPU: two gloo processes each run the oracle on their frame shard; the gathered per-frame
digests must equal those of an unsharded run (the N>1 path of bench.py / hevcasm_b200.shard, without GPUs)."""
import os
import subprocess
import sys

import pytest

from hevcasm_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from hevcasm_b200 import shard, synth
from hevcasm_b200.abi import HEVCASM_RECT
from oracle import binding
from oracle.binding import ptr

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
NF, W, H = 5, 128, 64
first, last = shard.frame_range(NF, rank, world)
oracle = binding.oracle()
digs = []
for f in range(first, last):   # every rank regenerates only its own frames (seeded per frame)
    src = synth.random_planes(1000 + f, 1, W, H, 16)
    ref = synth.random_planes(2000 + f, 1, W, H, 16)
    out = np.zeros((W // 16) * (H // 16) * 64, np.int32)
    oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, HEVCASM_RECT(16, 16), -4, -4, 8, 8, 1,
               src.frame_stride, ref.frame_stride, ptr(out))
    digs.append(shard.frame_digest(out))
all_digs = shard.gather_frame_digests(digs, first, NF, dist)
dist.barrier()
if rank == 0:
    print(json.dumps(all_digs))
dist.destroy_process_group()
"""


def test_frame_range_partitions_the_batch():
    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard.frame_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_run_matches_unsharded(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29600 + world + os.getpid() % 500), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("[")][-1]
    import json
    digs = json.loads(line)
    assert len(digs) == 5 and len(set(digs)) == 5
    if not hasattr(test_sharded_run_matches_unsharded, "ref"):
        test_sharded_run_matches_unsharded.ref = digs
    assert digs == test_sharded_run_matches_unsharded.ref
