"""GPU: the feed kernel (pndf_feed_batch) against the REAL reference loader's outputs (golden from PoseData.__getitem__,
tests/golden/make_data_golden.py) given the indices the reference drew, and the statistics of its own in-kernel sampling."""
import os

import numpy as np
import pytest
import torch

from posendf_b200.data import ResidentPoseData

from golden.make_data_golden import N_AMASS, N_FILES, NUM_PTS, write_files

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posedata.npz")


@pytest.mark.parametrize("flip", [False, True])
def test_batch_equals_reference_getitem_bit_for_bit(tmp_path, flip):
    data, amass = write_files(str(tmp_path))
    z = np.load(GOLD)
    feed = ResidentPoseData(data, amass, batch_size=N_FILES, num_pts=NUM_PTS, flip=flip, device="cuda")
    tags = [f"f{int(flip)}_i{i}" for i in range(N_FILES)]
    rows = np.stack([z[t + "_rows"] for t in tags])
    arows = np.stack([z[t + "_amass_rows"] for t in tags])
    ai = [int(z[t + "_amass_idx"]) for t in tags]
    out = feed.batch(list(range(N_FILES)), ai, rows=rows, amass_rows=arows)      # the whole DataLoader batch in one launch
    torch.cuda.synchronize()
    for i, t in enumerate(tags):
        assert np.array_equal(out["pose"][i].cpu().numpy(), z[t + "_pose"])
        assert np.array_equal(out["man_poses"][i].cpu().numpy(), z[t + "_man_poses"])
        assert np.allclose(out["dist"][i].cpu().numpy(), z[t + "_dist"], rtol=1e-6, atol=0)
    if flip:
        fixed = ResidentPoseData(data, amass, batch_size=N_FILES, num_pts=NUM_PTS, flip=True, device="cuda", fix_flip_bug=True)
        o2 = fixed.batch(list(range(N_FILES)), ai, rows=rows, amass_rows=arows)
        assert (o2["man_poses"][..., 0] >= 0).all() and not torch.equal(o2["man_poses"], o2["pose"])
        assert torch.equal(o2["pose"], out["pose"])


def test_in_kernel_sampling_is_uniform_with_replacement_and_reproducible(tmp_path):
    data, amass = write_files(str(tmp_path))
    feed = ResidentPoseData(data, amass, batch_size=2, num_pts=20000, device="cuda", seed=5)
    a = feed.batch([1, 3], [0, 2], seed=77)
    b = feed.batch([1, 3], [0, 2], seed=77)
    c = feed.batch([1, 3], [0, 2], seed=78)
    assert all(torch.equal(a[k], b[k]) for k in a) and not torch.equal(a["pose"], c["pose"])
    # every sampled pose is a row of the item's own file; all rows get hit about equally often (with replacement)
    for item, f in enumerate((1, 3)):
        src = torch.from_numpy(np.load(data[f])["pose"].reshape(-1, 84)).cuda()
        got = a["pose"][item].reshape(-1, 84)
        match = (got[:, None, :4] == src[None, :, :4]).all(-1)          # first joint identifies the row (synthetic data: unique)
        assert match.any(1).all()
        counts = match.float().sum(0)
        exp = 20000 / len(src)
        assert counts.min().item() > 0.5 * exp and counts.max().item() < 1.6 * exp
    batches = list(feed)
    assert len(batches) == len(feed) and batches[0]["pose"].shape == (2, 20000, 21, 4)
