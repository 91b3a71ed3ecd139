"""Pose-batch sharding across GPUs (one process per GPU, torch.distributed).

Every pose is independent in forward, gradient, projection and the denoise prior term, so the only
collective on the path is ONE all-gather of the projected poses (and distances) at the end of a run
(SURVEY 8e).  Shards are contiguous slices aligned to the kernel's 32-pose tile so that a sharded run is
bit-identical to the single-GPU run.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE = 32


def shard_bounds(n: int, world: int, rank: int, align: int = TILE):
    """[lo, hi) of `rank`'s contiguous slice of n poses; slices are `align`-aligned, cover [0,n) exactly once,
    and differ by at most one tile."""
    tiles = (n + align - 1) // align
    base, extra = divmod(tiles, world)
    lo_t = rank * base + min(rank, extra)
    hi_t = lo_t + base + (1 if rank < extra else 0)
    return min(lo_t * align, n), min(hi_t * align, n)


def all_gather_ragged(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all-gather of the per-rank slices produced by shard_bounds (sizes may differ by one tile): pad to the
    largest slice, one all_gather_into_tensor, strip the padding.  Works with nccl (CUDA) and gloo (CPU)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    mx = max(counts)
    tail = local.shape[1:]
    padded = local
    if local.shape[0] != mx:
        padded = torch.zeros((mx, *tail), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * mx, *tail), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "gloo":
        chunks = list(out.chunk(world))
        dist.all_gather(chunks, padded.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


def project_sharded(net, poses: torch.Tensor, steps: int = 10, renorm: bool = False, group=None):
    """SamplePose.project over a batch sharded across the ranks of `group`: every rank passes the FULL batch
    (or at least its own slice region), projects its slice with the fused kernel and all ranks receive the full
    projected batch + distances.  One NCCL all-gather at the end, no per-step communication."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = poses.reshape(-1, 21, 4).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    x, d = net.project(poses.reshape(-1, 21, 4)[lo:hi], steps=steps, renorm=renorm)
    return all_gather_ragged(x, n, group), all_gather_ragged(d, n, group)


def allreduce_gradients(net, group=None):
    """Data-parallel training (config 5): average the parameter gradients of all ranks after backward() with ONE
    all-reduce of the flattened 1 365 565 fp32 values (5.46 MB; losses are means over equal shards, so averaging the
    per-rank gradients reproduces the single-process step)."""
    params = [p for p in net.parameters() if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
