"""CPU: host side of the device-resident data feed (posendf_b200/data.py) -- file tables, offsets, the epoch plan (shuffle,
drop_last, one AMASS file per item) -- and the golden of the REAL reference loader (tests/golden/make_data_golden.py lifts
PoseData out of /root/reference/model/load_data.py with ast) against an independent numpy restatement.  The gather itself is
the CUDA feed kernel: tests/test_data_feed.py (-m gpu)."""
import os

import numpy as np
import pytest

from posendf_b200.data import ResidentPoseData

from golden.make_data_golden import N_AMASS, N_FILES, NUM_PTS, write_files

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posedata.npz")


def test_tables_offsets_and_epoch_plan(tmp_path):
    data, amass = write_files(str(tmp_path))
    feed = ResidentPoseData(data, amass, batch_size=2, num_pts=50, device="cpu", seed=3)
    assert feed.n_files == N_FILES and feed.n_amass == N_AMASS
    pose, dist, am = feed._host
    for i, f in enumerate(data):
        z = np.load(f)
        lo, hi = feed.file_off_host[i], feed.file_off_host[i + 1]
        assert np.array_equal(pose[lo:hi], z["pose"].reshape(-1, 84)) and np.array_equal(dist[lo:hi], z["dist"])
    for i, f in enumerate(amass):
        lo, hi = feed.amass_off_host[i], feed.amass_off_host[i + 1]
        assert np.array_equal(am[lo:hi], np.load(f)["pose"].reshape(-1, 84))
    plan = feed.plan_epoch()
    assert len(plan) == len(feed) == N_FILES // 2                     # drop_last
    seen = np.concatenate([p[0] for p in plan])
    assert len(set(seen.tolist())) == len(seen) and set(seen.tolist()) <= set(range(N_FILES))      # a shuffle without repeats
    for files, am_idx, seed in plan:
        assert files.dtype == np.int32 and am_idx.dtype == np.int32 and len(files) == len(am_idx) == 2
        assert (am_idx >= 0).all() and (am_idx < N_AMASS).all() and 0 <= seed < 2 ** 63
    assert any(not np.array_equal(a[0], b[0]) for a, b in zip(plan, feed.plan_epoch())) or N_FILES < 3      # reshuffled


def test_no_cpu_gather():
    import tempfile
    with tempfile.TemporaryDirectory() as root:
        data, amass = write_files(root)
        feed = ResidentPoseData(data, amass, batch_size=2, num_pts=8, device="cpu")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            feed.batch([0, 1], [0, 0])


def test_reference_golden_matches_numpy_restatement(tmp_path):
    """pins the meaning of the golden: PoseData.__getitem__ == rows gather, flip, mean of 5, AMASS gather (+ the flip quirk)"""
    data, amass = write_files(str(tmp_path))
    z = np.load(GOLD)
    assert int(z["NUM_PTS"]) == NUM_PTS
    for flip in (0, 1):
        for idx in range(N_FILES):
            t = f"f{flip}_i{idx}"
            src = np.load(data[idx])
            rows, ai, arows = z[t + "_rows"], int(z[t + "_amass_idx"]), z[t + "_amass_rows"]
            pose = src["pose"][rows]
            if flip:
                pose = np.where(pose[:, :, :1] < 0, -pose, pose)
            man = pose if flip else np.load(amass[ai])["pose"][arows]
            assert np.array_equal(z[t + "_pose"], pose) and np.array_equal(z[t + "_man_poses"], man)
            assert np.allclose(z[t + "_dist"], src["dist"][rows].mean(axis=1), rtol=1e-6)
