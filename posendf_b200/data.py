"""Device-resident training-data feed, mirroring the batches of the reference's `PoseData`
(/root/reference/model/load_data.py:18-86) without its 30 DataLoader worker processes.

The reference opens one `.npz` per item (`{'pose' (N,21,4), 'dist' (N,5), 'nn_pose'}`, written by
data/prepare_traindata.py:173), draws `num_pts` random rows with replacement, averages the 5 kNN distances, and pairs
them with `num_pts` random rows of ONE randomly chosen AMASS file; the DataLoader stacks `batch_size` such items
(`shuffle=True, drop_last=True`).  At 262 144 poses per step the per-item `np.load` + fancy indexing on the host is the
bottleneck, so here every file is read ONCE, concatenated into three tables in HBM (AMASS-scale data is a few GB, HBM has
180), and a whole batch is ONE launch of the library's feed kernel (pndf_feed_batch, csrc/pndf_feed.cuh: in-kernel
counter-based row sampling, 336-byte row gather, quaternion flip, mean of the 5 distances).  Host code only plans the epoch
(file order, which AMASS file per item); there is no CPU / torch implementation of the gather.

Faithful quirks (kept, because a drop-in must feed the trainer the same distribution):
  * rows are sampled WITH replacement (`np.random.randint`), load_data.py:49,60;
  * `dist` is the MEAN of the 5 stored neighbour distances, load_data.py:53;
  * with `flip=True` the reference overwrites the manifold poses with the flipped NOISY poses
    (`amass_poses, _ = quat_flip(poses)`, load_data.py:63) -- reproduced unless `fix_flip_bug=True`.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from . import _lib


class ResidentPoseData:
    """Iterable of batches `{'pose': (b, n, 21, 4), 'dist': (b, n), 'man_poses': (b, n, 21, 4)}` on `device` (CUDA)."""

    def __init__(self, data_files, amass_files, batch_size=4, num_pts=5000, flip=False, device="cuda", seed=None,
                 fix_flip_bug=False):
        if len(data_files) == 0 or len(amass_files) == 0:
            raise ValueError("ResidentPoseData needs at least one data file and one AMASS file")
        self.device = torch.device(device)
        self.batch_size, self.num_pts, self.flip, self.fix_flip_bug = int(batch_size), int(num_pts), bool(flip), bool(fix_flip_bug)
        self.host_gen = np.random.default_rng(seed)       # host-side choices: file order, AMASS file per item, kernel seeds
        pose, dist, off = [], [], [0]
        for f in data_files:
            z = np.load(f)
            p = np.asarray(z["pose"], dtype=np.float32).reshape(-1, 84)
            d = np.asarray(z["dist"], dtype=np.float32).reshape(len(p), 5)
            pose.append(p); dist.append(d); off.append(off[-1] + len(p))
        am, aoff = [], [0]
        for f in amass_files:
            p = np.asarray(np.load(f)["pose"], dtype=np.float32).reshape(-1, 84)
            am.append(p); aoff.append(aoff[-1] + len(p))
        self.n_files, self.n_amass = len(data_files), len(amass_files)
        self.file_off_host, self.amass_off_host = np.asarray(off, dtype=np.int64), np.asarray(aoff, dtype=np.int64)
        self._host = (np.concatenate(pose), np.concatenate(dist), np.concatenate(am))
        self._dev = None

    @classmethod
    def from_dirs(cls, data_path, amass_dir, splits, **kw):
        """same file discovery as the reference (load_data.py:27-32): `<data_path>/<dataset>/*000.npz`, `<amass_dir>/<dataset>/*.npz`"""
        data = [f for f in sorted(glob.glob(os.path.join(data_path, "*", "*000.npz"))) if f.split("/")[-2] in splits]
        amass = [f for f in sorted(glob.glob(os.path.join(amass_dir, "*", "*.npz"))) if f.split("/")[-2] in splits]
        return cls(data, amass, **kw)

    def __len__(self):
        return self.n_files // self.batch_size          # drop_last=True

    # ------------------------------------------------------------------ host: epoch plan
    def plan_epoch(self):
        """[(item_files (b,), item_amass (b,), kernel seed)] for one epoch: shuffle=True, drop_last=True (load_data.py:75-77);
        every item gets ONE random AMASS file (load_data.py:57)"""
        order = self.host_gen.permutation(self.n_files)
        plan = []
        for b in range(len(self)):
            files = order[b * self.batch_size:(b + 1) * self.batch_size].astype(np.int32)
            amass = self.host_gen.integers(0, self.n_amass, self.batch_size).astype(np.int32)
            plan.append((files, amass, int(self.host_gen.integers(0, 2 ** 63 - 1))))
        return plan

    # ------------------------------------------------------------------ device
    def _tables(self):
        if self._dev is None:
            if self.device.type != "cuda":
                raise RuntimeError("ResidentPoseData: batches are assembled by the CUDA feed kernel (pndf_feed_batch); there is "
                                   f"no CPU fallback (device = {self.device})")
            up = lambda a: torch.from_numpy(a).to(self.device)      # noqa: E731
            self._dev = tuple(up(a) for a in self._host) + (up(self.file_off_host), up(self.amass_off_host))
            self._host = None
        return self._dev

    def batch(self, item_files, item_amass, seed=0, rows=None, amass_rows=None):
        """one DataLoader batch in ONE launch; rows / amass_rows (b, num_pts) int64 inject the row indices (tests)"""
        pose_t, dist_t, amass_t, foff, aoff = self._tables()
        lib = _lib.load()
        b, n = len(item_files), self.num_pts
        dev = self.device
        fi = torch.as_tensor(np.asarray(item_files, dtype=np.int32)).to(dev)
        ai = torch.as_tensor(np.asarray(item_amass, dtype=np.int32)).to(dev)
        pose = torch.empty(b, n, 21, 4, device=dev, dtype=torch.float32)
        dist = torch.empty(b, n, device=dev, dtype=torch.float32)
        man = torch.empty(b, n, 21, 4, device=dev, dtype=torch.float32)
        r = None if rows is None else torch.as_tensor(np.asarray(rows, dtype=np.int64)).to(dev).contiguous()
        ar = None if amass_rows is None else torch.as_tensor(np.asarray(amass_rows, dtype=np.int64)).to(dev).contiguous()
        _lib.check(lib.pndf_feed_batch(dev.index if dev.index is not None else torch.cuda.current_device(), pose_t.data_ptr(),
                                       dist_t.data_ptr(), foff.data_ptr(), amass_t.data_ptr(), aoff.data_ptr(), fi.data_ptr(),
                                       ai.data_ptr(), b, n, int(self.flip), int(self.fix_flip_bug), int(seed) & (2 ** 64 - 1),
                                       None if r is None else r.data_ptr(), None if ar is None else ar.data_ptr(),
                                       pose.data_ptr(), dist.data_ptr(), man.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return {"pose": pose, "dist": dist, "man_poses": man}

    def __iter__(self):
        for files, amass, seed in self.plan_epoch():
            yield self.batch(files, amass, seed)
