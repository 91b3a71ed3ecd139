"""Pose-batch sharding across GPUs (one process per GPU, torch.distributed).

Every pose is independent in forward, gradient, projection and the denoise prior term, so the only
collective on the path is ONE all-gather of the projected poses (and distances) at the end of a run
(SURVEY 8e).  Shards are contiguous slices aligned to the kernel's 32-pose tile so that a sharded run is
bit-identical to the single-GPU run.
"""
from __future__ import annotations

import ctypes as C

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
    eng = net.engine()
    eng.set_tile_policy(eng.tile_for_batch(n))       # the tiling of the WHOLE batch: sharded == unsharded bit for bit
    try:
        x, d = net.project(poses.reshape(-1, 21, 4)[lo:hi], steps=steps, renorm=renorm)
    finally:
        eng.set_tile_policy(0)
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


# ------------------------------------------------------------------------------------------------ gathered projection runs
class _RawCuda:
    """zero-copy torch view of library-owned device memory (__cuda_array_interface__)"""

    def __init__(self, ptr, nfloat, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (nfloat,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _world_rank(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class LocalGather:
    """world size 1: the gathered buffer is the local buffer"""
    kind = "none (1 GPU)"

    def __init__(self, n_local, device):
        self.n, self.device = n_local, device
        self.poses = torch.empty(n_local, 21, 4, device=device, dtype=torch.float32)
        self.dist = torch.empty(n_local, 1, device=device, dtype=torch.float32)

    def local_view(self):
        return self.poses

    def project_and_gather(self, eng, steps=1, renorm=False):
        import posendf_b200._lib as _lib
        _lib.check(eng.lib.pndf_project(eng._h, self.poses.data_ptr(), self.n, int(steps), int(renorm), self.dist.data_ptr(),
                                        torch.cuda.current_stream(self.device).cuda_stream))
        return self.dist


class NcclGather:
    """fallback: the projection kernel, then ONE all_gather_into_tensor of the projected poses (and one of the distances)
    on the same stream -- equal per-rank slices of n_local poses"""
    kind = "NCCL all_gather_into_tensor after the kernel"

    def __init__(self, n_local, device, group=None):
        self.world, self.rank = _world_rank(group)
        self.n, self.device, self.group = n_local, device, group
        self.poses = torch.empty(self.world * n_local, 21, 4, device=device, dtype=torch.float32)
        self.dist = torch.empty(self.world * n_local, 1, device=device, dtype=torch.float32)

    def local_view(self):
        return self.poses[self.rank * self.n:(self.rank + 1) * self.n]

    def project_and_gather(self, eng, steps=1, renorm=False):
        import posendf_b200._lib as _lib
        x = self.local_view()
        d = self.dist[self.rank * self.n:(self.rank + 1) * self.n]
        eng.set_tile_policy(eng.tile_for_batch(self.world * self.n))      # tiling of the whole batch (bit-identical to unsharded)
        try:
            _lib.check(eng.lib.pndf_project(eng._h, x.data_ptr(), self.n, int(steps), int(renorm), d.data_ptr(),
                                            torch.cuda.current_stream(self.device).cuda_stream))
        finally:
            eng.set_tile_policy(0)
        dist.all_gather_into_tensor(self.poses, x, group=self.group)
        dist.all_gather_into_tensor(self.dist, d, group=self.group)
        return d


class PeerGather:
    """The gather fused into the projection kernel (pndf_project_gather): every rank owns a cudaIpc-shared buffer
    [world * n_local poses | world * n_local distances | flags]; the kernel's write-back stores each projected tile into
    all peers' buffers over NVLink while the following tiles compute, and a peer-memory barrier (pndf_peer_barrier)
    closes the step.  No NCCL on the data path; torch.distributed only carries the 64-byte IPC handles once."""
    kind = "fused into the kernel: peer (NVLink / cudaIpc) stores from the write-back + peer-memory barrier"

    def __init__(self, n_local, device, group=None):
        import posendf_b200._lib as _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.world, self.rank = _world_rank(group)
        if not (2 <= self.world <= 8):
            raise RuntimeError("PeerGather needs 2..8 ranks on one node")
        self.n, self.device, self.group = n_local, device, group
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        W, n = self.world, n_local
        self.pose_floats, self.dist_floats = W * n * 84, ((W * n + 3) // 4) * 4
        nbytes = (self.pose_floats + self.dist_floats) * 4 + 256
        self._flag_off = (self.pose_floats + self.dist_floats) * 4
        base, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(self.lib.pndf_peer_alloc(self.dev_index, nbytes, C.byref(base), handle))
        self._base = base.value
        self._opened = []
        # ship the handles (NCCL moves CUDA tensors, gloo CPU tensors)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device if dist.get_backend(group) == "nccl" else "cpu")
        allh = [torch.empty_like(mine) for _ in range(W)]
        dist.all_gather(allh, mine, group=group)
        self._bases = [0] * W
        for r in range(W):
            if r == self.rank:
                self._bases[r] = self._base
                continue
            hb = (C.c_ubyte * 64)(*allh[r].cpu().tolist())
            ptr = C.c_void_p()
            _lib.check(self.lib.pndf_peer_open(self.dev_index, hb, C.byref(ptr)))
            self._bases[r] = ptr.value
            self._opened.append(ptr.value)
        self.poses = torch.as_tensor(_RawCuda(self._base, self.pose_floats), device=device).view(W * n, 21, 4)
        self.dist = torch.as_tensor(_RawCuda(self._base + self.pose_floats * 4, W * n), device=device).view(W * n, 1)
        peers = [r for r in range(W) if r != self.rank]
        PP = C.c_void_p * max(1, len(peers))
        self._peer_pose = PP(*[self._bases[r] + self.rank * n * 84 * 4 for r in peers])
        self._peer_dist = PP(*[self._bases[r] + self.pose_floats * 4 + self.rank * n * 4 for r in peers])
        self._flags = (C.c_void_p * W)(*[self._bases[r] + self._flag_off for r in range(W)])
        self._npeers = len(peers)
        self._epoch = 0
        dist.barrier(group=group)         # every rank has mapped every buffer before the first store

    def local_view(self):
        return self.poses[self.rank * self.n:(self.rank + 1) * self.n]

    def barrier(self):
        self._epoch += 1
        self._lib.check(self.lib.pndf_peer_barrier(self.dev_index, self._flags, self.world, self.rank, self._epoch,
                                                   torch.cuda.current_stream(self.device).cuda_stream))

    def project_and_gather(self, eng, steps=1, renorm=False):
        x = self.local_view()
        d = self.dist[self.rank * self.n:(self.rank + 1) * self.n]
        st = torch.cuda.current_stream(self.device).cuda_stream
        eng.set_tile_policy(eng.tile_for_batch(self.world * self.n))      # tiling of the whole batch (bit-identical to unsharded)
        try:
            self._lib.check(self.lib.pndf_project_gather(eng._h, x.data_ptr(), self.n, int(steps), int(renorm), d.data_ptr(),
                                                         self._peer_pose, self._peer_dist, self._npeers, st))
        finally:
            eng.set_tile_policy(0)
        self.barrier()
        return d

    def timed_out(self) -> bool:
        """True if a peer barrier gave up waiting (a rank died); synchronises"""
        flags = torch.as_tensor(_RawCuda(self._base + self._flag_off, self.world + 1, "<u4"), device=self.device)
        return bool(flags[self.world].item())

    def close(self):
        if getattr(self, "_base", None):
            torch.cuda.synchronize(self.device)
            self.poses = self.dist = None
            for p in self._opened:
                self.lib.pndf_peer_close(self.dev_index, p)
            self.lib.pndf_peer_free(self.dev_index, self._base)
            self._base = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_gather(n_local, device, prefer="auto", group=None):
    """gathered-projection helper for this process group: PeerGather where peer memory can be mapped (all ranks must agree),
    NcclGather otherwise; LocalGather for a single process."""
    world, _ = _world_rank(group)
    if world == 1:
        return LocalGather(n_local, device)
    if prefer == "nccl":
        return NcclGather(n_local, device, group)
    g, ok = None, 1
    try:
        g = PeerGather(n_local, device, group)
    except Exception as e:      # noqa: BLE001
        ok, err = 0, e
    flag = torch.tensor([ok], device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if flag.item() == 1:
        return g
    if prefer == "peer":
        raise RuntimeError(f"peer gather unavailable on some rank ({'' if ok else err})")
    if g is not None:
        g.close()
    return NcclGather(n_local, device, group)


# ------------------------------------------------------------------------------------------------ data-parallel trainer step
class DataParallelStep:
    """One trainer step of model/train_posendf.py:93-99 on every rank of `group` (config C5): zero_grad, fused
    train-mode forward (dist + manifold + Eikonal), loss = sum_k w_k L_k, backward, ONE all-reduce of the flat gradient
    vector (mean over ranks = the single-process gradient of the global batch), Adam(lr, weight_decay)
    (model/train_posendf.py:30).  Uses posendf_b200.optim.FusedAdam (one kernel: moments, update and the re-packed slab
    stream of the engine) on the module's flat parameter / gradient buffers."""

    def __init__(self, net, lr=1e-5, weight_decay=1e-4, weights=(1.0, 1.0, 1.0), group=None):
        from .optim import FusedAdam
        self.net, self.group = net, group
        self.w_dist, self.w_man, self.w_eik = (float(w) for w in weights)
        self.world, _ = _world_rank(group)
        self.optim = FusedAdam(net, lr=lr, weight_decay=weight_decay)
        self.kind = self.optim.kind

    def step(self, pose, dist_gt, man_poses):
        self.optim.zero_grad()
        _, ld = self.net(pose, dist_gt, man_poses, train=True, eikonal=self.w_eik)
        w = {"dist": self.w_dist, "man_loss": self.w_man, "eikonal": self.w_eik}
        loss = sum(w[k] * v for k, v in ld.items())
        loss.backward()
        if self.world > 1:
            # ONE all-reduce of the flat 1 365 565-float gradient vector; the 1/world of the mean is applied inside the
            # optimizer kernel
            dist.all_reduce(self.optim.flat_grad(), op=dist.ReduceOp.SUM, group=self.group)
            self.optim.grad_scale = 1.0 / self.world
        self.optim.step()
        return ld
