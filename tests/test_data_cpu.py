"""CPU: the device-resident data feed reproduces `PoseData.__getitem__` (restated in numpy from
/root/reference/model/load_data.py:43-71) when given the same indices."""
import numpy as np
import torch

from posendf_b200 import synth
from posendf_b200.data import ResidentPoseData, quat_flip


def _write_files(tmp_path, n_files=5, n_amass=3):
    data, amass = [], []
    for i in range(n_files):
        d = tmp_path / f"ds{i % 2}"
        d.mkdir(exist_ok=True)
        n = 200 + 17 * i
        pose = synth.make_poses(10 + i, n, kind="noisy", sigma=0.3)
        pose[::3] *= -1                                           # some negative real parts for the flip
        f = d / f"part{i}_000.npz"
        np.savez(f, pose=pose, dist=np.abs(synth.normal(20 + i, n * 5)).reshape(n, 5).astype(np.float32), nn_pose=pose[:, None])
        data.append(str(f))
    for i in range(n_amass):
        d = tmp_path / f"am{i}"
        d.mkdir(exist_ok=True)
        pose = synth.make_poses(50 + i, 150 + i)
        pose[1::4] *= -1
        f = d / f"seq{i}.npz"
        np.savez(f, pose=pose)
        amass.append(str(f))
    return data, amass


def _reference_item(data_file, amass_files, rows, amass_idx, amass_rows, flip):
    z = np.load(data_file)
    poses = z["pose"][rows]
    def qflip(p):
        q = np.copy(p); neg = p[:, :, 0] < 0; q[np.where(neg)] = -q[np.where(neg)]; return q
    if flip:
        poses = qflip(poses)
    dist = np.mean(z["dist"][rows], axis=1)
    am = np.load(amass_files[amass_idx])["pose"][amass_rows]
    if flip:
        am = qflip(poses)                                          # the reference's own behaviour (load_data.py:63)
    return poses.astype(np.float32), dist.astype(np.float32), am.astype(np.float32)


def test_item_matches_reference_semantics(tmp_path):
    data, amass = _write_files(tmp_path)
    rng = np.random.default_rng(0)
    for flip in (False, True):
        feed = ResidentPoseData(data, amass, batch_size=2, num_pts=64, flip=flip, device="cpu", seed=1)
        for idx in range(len(data)):
            rows = rng.integers(0, len(feed.pose[idx]), 64)
            ai = int(rng.integers(0, len(amass)))
            arows = rng.integers(0, len(feed.amass[ai]), 64)
            it = feed.item(idx, torch.from_numpy(rows), ai, torch.from_numpy(arows))
            p, d, m = _reference_item(data[idx], amass, rows, ai, arows, flip)
            assert np.array_equal(it["pose"].numpy(), p) and np.allclose(it["dist"].numpy(), d, rtol=1e-6) and np.array_equal(it["man_poses"].numpy(), m)
    fixed = ResidentPoseData(data, amass, num_pts=32, flip=True, device="cpu", fix_flip_bug=True).item(0)
    assert (fixed["man_poses"][..., 0] >= 0).all() and not torch.equal(fixed["man_poses"], fixed["pose"])
    assert torch.equal(quat_flip(torch.tensor([[-1.0, 2, 3, 4], [1.0, -2, 3, 4]])), torch.tensor([[1.0, -2, -3, -4], [1.0, -2, 3, 4]]))


def test_epoch_shapes_shuffle_and_drop_last(tmp_path):
    data, amass = _write_files(tmp_path)
    feed = ResidentPoseData(data, amass, batch_size=2, num_pts=50, device="cpu", seed=3)
    batches = list(feed)
    assert len(batches) == len(feed) == 2                          # 5 files, batch 2, drop_last
    for b in batches:
        assert b["pose"].shape == (2, 50, 21, 4) and b["dist"].shape == (2, 50) and b["man_poses"].shape == (2, 50, 21, 4)
        assert b["pose"].dtype == torch.float32 and torch.isfinite(b["dist"]).all()
