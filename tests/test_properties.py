"""Property tests: CPU ones exercise the oracle, GPU ones the fused kernel through the C ABI."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import assert_grad_parity, rel_err
from oracle import posendf_numpy as onp
from posendf_b200 import synth

PARAMS64 = {k: v.astype(np.float64) for k, v in synth.make_params(1).items()}


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10_000), scale=st.floats(0.1, 10.0), act=st.sampled_from(["relu", "lrelu", "softplus"]))
def test_oracle_invariances(seed, scale, act):
    """d >= 0; the column normalisation makes d invariant to a positive rescaling of the whole pose tensor and the
    gradient scale as 1/scale; swapping the inputs of two leaf joints changes exactly the features the table predicts."""
    cfg = onp.default_cfg(enc_act=act, df_act=act, enc_beta=30.0, df_beta=30.0)
    x = synth.make_poses(seed, 4, kind="raw", dtype=np.float64)
    d, g = onp.forward_grad(PARAMS64, x, cfg)
    assert (d >= 0).all()
    d2, g2 = onp.forward_grad(PARAMS64, x * scale, cfg)
    assert np.allclose(d2, d, rtol=1e-10, atol=1e-12)
    assert np.allclose(g2 * scale, g, rtol=1e-8, atol=1e-12)
    # joints 10 and 15 are leaves with parents 8 and 13: perturbing joint 10 may only move features of joint 10
    q, _ = onp.normalise_columns(x)
    z, _ = onp.encoder_forward(PARAMS64, q, cfg)
    q2 = q.copy(); q2[:, 10, :] += 0.1
    z2, _ = onp.encoder_forward(PARAMS64, q2, cfg)
    changed = np.abs(z2 - z).reshape(len(x), 21, 6).max(axis=(0, 2)) > 0
    assert not changed[[i for i in range(21) if i != 10]].any()
    # joint 9 feeds joints 11,12,13 and everything below 12/13
    q3 = q.copy(); q3[:, 9, :] += 0.1
    z3, _ = onp.encoder_forward(PARAMS64, q3, cfg)
    changed = np.abs(z3 - z).reshape(len(x), 21, 6).max(axis=(0, 2)) > 0
    assert not changed[[0, 1, 2, 3, 4, 5, 6, 7, 8, 10]].any()


def test_oracle_sign_flip_of_one_component_column():
    """F.normalize(dim=1) + the first Linear see q, not |q|: flipping the sign of a whole component column is NOT an
    invariance (unlike q -> -q for a true rotation), the oracle must reproduce that reference behaviour."""
    cfg = onp.default_cfg()
    x = synth.make_poses(3, 6, dtype=np.float64)
    d = onp.forward(PARAMS64, x, cfg)
    x2 = x.copy(); x2[:, :, 1] *= -1
    assert np.abs(onp.forward(PARAMS64, x2, cfg) - d).max() > 1e-6


@pytest.mark.gpu
@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10_000), B=st.integers(1, 200), act=st.sampled_from(["relu", "lrelu", "softplus"]),
       beta=st.sampled_from([5.0, 30.0, 100.0]), kind=st.sampled_from(["randn", "rand", "noisy", "raw"]),
       wseed=st.integers(1, 9))
def test_kernel_matches_oracle_on_random_configs(seed, B, act, beta, kind, wseed):
    from posendf_b200.engine import Engine
    params = synth.make_params(wseed)
    eng = Engine(device=0, enc_act=act, df_act=act, enc_beta=beta, df_beta=beta)
    eng.set_weights_flat(synth.flatten_params(params))
    cfg = onp.default_cfg(enc_act=act, df_act=act, enc_beta=beta, df_beta=beta)
    poses = synth.make_poses(seed, B, kind=kind)
    d, g = eng.forward_grad(torch.from_numpy(poses).cuda())
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    dref, gref = onp.forward_grad(p64, poses.astype(np.float64), cfg)
    assert np.max(rel_err(d.cpu().numpy(), dref)) < 1e-5
    assert (d >= 0).all()
    if B >= 50:
        assert_grad_parity(g.cpu().numpy(), gref, outlier_frac=0.04)
    else:
        e = np.linalg.norm((g.cpu().numpy() - gref).reshape(B, -1), axis=1) / np.linalg.norm(gref.reshape(B, -1), axis=1)
        assert np.median(e) < 1e-5
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("act", ["lrelu", "softplus"])
def test_kernel_tree_table_perturbing_a_joint_moves_exactly_the_predicted_feature_rows(act):
    """the KERNEL's copy of the kinematic parent table, read back through the debug dump of the encoder
    output z0 (rows [0, 126) = joint j's six features at rows 6j .. 6j+5).  Perturbing the quaternion of joint j must change
    the features of j and of every descendant of j in the reference's table (model/network/net_utils.py:44-50, golden
    parents.npz) -- and of nothing else.  The perturbation keeps the four column norms of F.normalize(dim=1) fixed (the same
    joint's quaternion is replaced by another with the same squared components), so only the tree couples joints."""
    import os
    from conftest import GOLDEN_DIR
    from posendf_b200.engine import Engine
    parents = np.load(os.path.join(GOLDEN_DIR, "parents.npz"))["parents"].tolist()
    desc = {j: {j} for j in range(21)}
    for i in range(21):                       # index order is topological
        p = parents[i]
        while p >= 0:
            desc[p].add(i)
            p = parents[p]
    eng = Engine(device=0, enc_act=act, df_act=act)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
    x = synth.make_poses(17, 32)
    base = eng.forward_grad_debug(torch.from_numpy(x).cuda())[2][:, :126].cpu().numpy().reshape(32, 21, 6)
    for j in range(21):
        y = x.copy()
        y[:, j, :] = -y[:, j, :]              # same squares -> same column norms; the joint's normalised input flips sign
        z0 = eng.forward_grad_debug(torch.from_numpy(y).cuda())[2][:, :126].cpu().numpy().reshape(32, 21, 6)
        moved = {i for i in range(21) if not np.array_equal(z0[:, i], base[:, i])}
        assert moved <= desc[j], (j, sorted(moved - desc[j]))
        # the joint itself always moves; descendants move unless every unit on the path is dead for all 32 poses (relu-like)
        assert j in moved
        if act == "softplus":
            assert moved == desc[j], (j, sorted(desc[j] - moved))
