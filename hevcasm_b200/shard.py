"""Multi-GPU sharding of a frame batch: whole frames are dealt to ranks, every rank runs the same batched
entry points on its own frames, and nothing is exchanged on the data path.  The only communication is the bookkeeping
around it - a barrier for timing and an all-gather of per-frame output digests so that a sharded run can be compared with
an unsharded one - and goes through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import hashlib

import numpy as np


def frame_range(n_frames, rank, world):
    """contiguous, balanced slice [first, last) of the batch for `rank`: sizes differ by at most one frame"""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def frame_digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def gather_frame_digests(local, first, n_frames, dist=None):
    """local: digests of this rank's frames (frame first, first+1, ...).  Returns the list for all n_frames on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert len(local) == n_frames
        return list(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (first, list(local)))
    out = [None] * n_frames
    for f0, digs in parts:
        for i, d in enumerate(digs):
            out[f0 + i] = d
    assert all(d is not None for d in out), "frame shards do not cover the batch"
    return out
