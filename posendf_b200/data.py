"""Device-resident training-data feed, mirroring the batches of the reference's `PoseData`
(/root/reference/model/load_data.py:18-86) without its 30 DataLoader worker processes.

The reference opens one `.npz` per item (`{'pose' (N,21,4), 'dist' (N,5), 'nn_pose'}`, written by
data/prepare_traindata.py:173), draws `num_pts` random rows with replacement, averages the 5 kNN distances, and pairs
them with `num_pts` random rows of ONE randomly chosen AMASS file; the DataLoader stacks `batch_size` such items
(`shuffle=True, drop_last=True`).  At 262 144 poses per step the per-item `np.load` + fancy indexing on the host is the
bottleneck, so here every file is loaded ONCE into device memory (fp32; AMASS-scale data is a few GB, HBM has 180) and a
batch is two `randint` + gather kernels on the device.

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


def quat_flip(pose: torch.Tensor) -> torch.Tensor:
    """negate every quaternion whose real part is negative (load_data.py:12-16)"""
    return torch.where(pose[..., :1] < 0, -pose, pose)


class ResidentPoseData:
    """Iterable of batches `{'pose': (b, n, 21, 4), 'dist': (b, n), 'man_poses': (b, n, 21, 4)}` on `device`."""

    def __init__(self, data_files, amass_files, batch_size=4, num_pts=5000, flip=False, device="cuda", seed=None,
                 fix_flip_bug=False):
        if len(data_files) == 0 or len(amass_files) == 0:
            raise ValueError("ResidentPoseData needs at least one data file and one AMASS file")
        self.device = torch.device(device)
        self.batch_size, self.num_pts, self.flip, self.fix_flip_bug = int(batch_size), int(num_pts), bool(flip), bool(fix_flip_bug)
        self.gen = torch.Generator(device=self.device)
        self.host_gen = torch.Generator()       # host-side choices (file order, which AMASS file): no device sync per item
        if seed is not None:
            self.gen.manual_seed(int(seed))
            self.host_gen.manual_seed(int(seed))
        self.pose, self.dist = [], []
        for f in data_files:
            z = np.load(f)
            self.pose.append(torch.from_numpy(np.asarray(z["pose"], dtype=np.float32)).to(self.device))
            self.dist.append(torch.from_numpy(np.asarray(z["dist"], dtype=np.float32)).mean(dim=1).to(self.device))
        self.amass = [torch.from_numpy(np.asarray(np.load(f)["pose"], dtype=np.float32)).to(self.device) for f in amass_files]

    @classmethod
    def from_dirs(cls, data_path, amass_dir, splits, **kw):
        """same file discovery as the reference (load_data.py:27-32): `<data_path>/<dataset>/*000.npz`, `<amass_dir>/<dataset>/*.npz`"""
        data = [f for f in sorted(glob.glob(os.path.join(data_path, "*", "*000.npz"))) if f.split("/")[-2] in splits]
        amass = [f for f in sorted(glob.glob(os.path.join(amass_dir, "*", "*.npz"))) if f.split("/")[-2] in splits]
        return cls(data, amass, **kw)

    def __len__(self):
        return len(self.pose) // self.batch_size          # drop_last=True

    def _randint(self, high, n):
        return torch.randint(0, high, (n,), device=self.device, generator=self.gen)

    def item(self, idx, rows=None, amass_idx=None, amass_rows=None):
        """one `PoseData.__getitem__` (indices can be injected for testing)"""
        rows = self._randint(len(self.pose[idx]), self.num_pts) if rows is None else rows
        pose = self.pose[idx][rows]
        if self.flip:
            pose = quat_flip(pose)
        dist = self.dist[idx][rows]
        if amass_idx is None:
            amass_idx = int(torch.randint(0, len(self.amass), (1,), generator=self.host_gen))
        amass_rows = self._randint(len(self.amass[amass_idx]), self.num_pts) if amass_rows is None else amass_rows
        man = self.amass[amass_idx][amass_rows]
        if self.flip:
            man = quat_flip(man) if self.fix_flip_bug else pose       # reference bug, see module docstring
        return {"pose": pose, "dist": dist, "man_poses": man}

    def __iter__(self):
        order = torch.randperm(len(self.pose), generator=self.host_gen).tolist()                      # shuffle=True
        for b in range(len(self)):
            items = [self.item(i) for i in order[b * self.batch_size:(b + 1) * self.batch_size]]
            yield {k: torch.stack([it[k] for it in items]) for k in ("pose", "dist", "man_poses")}
