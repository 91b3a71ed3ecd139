"""Rotation formats on either side of the path (pytorch3d.transforms.axis_angle_to_quaternion / quaternion_to_axis_angle,
experiments/sample_poses.py:60,80, motion_denoise.py:81).  pytorch3d is not vendored in the reference and not installed:
PARITY UNPINNED at this boundary -- the restatements are checked against independent math (scipy's Rotation, round trips),
the kernels against the restatements."""
import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import posendf_numpy as onp
from posendf_b200 import synth


def _aa(n, seed=5, scale=1.0):
    return (synth.normal(seed, n * 3).reshape(n, 3) * scale).astype(np.float64)


def test_oracle_conversions_agree_with_scipy_rotations():
    aa = _aa(500)
    q = onp.axis_angle_to_quaternion(aa)                                   # (w, x, y, z)
    ref = Rotation.from_rotvec(aa).as_quat()                               # scipy: (x, y, z, w)
    ref = np.concatenate([ref[:, 3:], ref[:, :3]], axis=1)
    ref *= np.sign(np.sum(ref * q, axis=1, keepdims=True))                 # q and -q are the same rotation
    assert np.allclose(q, ref, atol=1e-12)
    back = onp.quaternion_to_axis_angle(q)
    assert np.allclose(back, aa, atol=1e-10)                               # all angles here are < pi
    # small-angle branches and the exact zero rotation
    tiny = _aa(50, scale=1e-9)
    tiny[0] = 0.0
    qt = onp.axis_angle_to_quaternion(tiny)
    assert np.all(np.isfinite(qt)) and np.allclose(qt[:, 0], 1.0) and np.allclose(qt[:, 1:], tiny / 2, atol=1e-24)
    assert np.allclose(onp.quaternion_to_axis_angle(qt), tiny, atol=1e-22)


@pytest.mark.gpu
def test_kernels_match_the_restatements_and_round_trip():
    from posendf_b200.engine import axis_angle_to_quaternion, quaternion_to_axis_angle
    aa = _aa(21 * 257, seed=8).astype(np.float32)
    aa[:21] *= 1e-9
    aa[3] = 0.0
    q = axis_angle_to_quaternion(torch.from_numpy(aa).cuda().reshape(257, 21, 3))
    assert q.shape == (257, 21, 4)
    assert np.allclose(q.cpu().numpy().reshape(-1, 4), onp.axis_angle_to_quaternion(aa.astype(np.float64)), atol=2e-7)
    back = quaternion_to_axis_angle(q)
    assert back.shape == (257, 21, 3)
    assert np.allclose(back.cpu().numpy().reshape(-1, 3), aa, rtol=2e-5, atol=2e-6)
    # arbitrary (non-unit, negative real part) quaternions, the input quaternion_to_axis_angle sees after projection steps
    qq = synth.make_poses(9, 300, kind="raw").reshape(-1, 4)
    got = quaternion_to_axis_angle(torch.from_numpy(qq).cuda()).cpu().numpy()
    assert np.allclose(got, onp.quaternion_to_axis_angle(qq.astype(np.float64)), rtol=2e-5, atol=2e-6)
