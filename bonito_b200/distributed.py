"""
Read-sharded multi-GPU execution: one process per GPU, chunks are independent from
`chunk()` to `stitch()` (`/root/reference/bonito/crf/basecall.py:63-77`), so the only collective on the path is one
broadcast of the parameters at start-up.  The reference itself has no multi-device support (single `--device`,
`bonito/cli/basecaller.py:177`).
"""

import torch
import torch.distributed as dist


def broadcast_parameters(model, src=0):
    """Rank `src`'s parameters and buffers to every rank (NCCL over NVLink on GPUs, gloo on CPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
    if hasattr(model, "invalidate_plan"):
        model.invalidate_plan()
    return model


def shard_reads(reads, rank=None, world=None):
    """Deal whole reads round-robin to ranks, so stitching stays rank-local and no result exchange is needed."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    for i, read in enumerate(reads):
        if i % world == rank:
            yield read


def gather_counts(n_samples, device="cpu"):
    """Sum of per-rank sample counts (the `samples per second` line of the CLI, bonito/cli/basecaller.py:160-164)."""
    if not (dist.is_available() and dist.is_initialized()):
        return n_samples
    t = torch.tensor([n_samples], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
