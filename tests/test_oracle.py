"""CPU: the oracle restatements (numpy closed form, torch autograd) against the golden vectors that
tests/golden/make_golden.py produced from the REAL reference modules."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, assert_grad_parity, assert_pose_parity, case_cfg, case_inputs, golden_case_names, load_golden, rel_err
from oracle import posendf_numpy as onp
from oracle import posendf_torch as otorch
from posendf_b200 import synth

CASES = golden_case_names()


def test_parent_table_bit_exact():
    ref = np.load(f"{GOLDEN_DIR}/parents.npz")["parents"]
    assert ref.dtype == np.int32 and ref.tolist() == list(onp.PARENTS) == list(synth.PARENTS) == list(otorch.PARENTS)


def test_column_normalise_matches_reference():
    z = np.load(f"{GOLDEN_DIR}/normalise.npz")
    x = synth.make_poses(77, 8, kind="raw")
    q32, _ = onp.normalise_columns(x)
    q64, _ = onp.normalise_columns(x.astype(np.float64))
    assert np.max(np.abs(q64 - z["q64"])) < 1e-15
    assert np.max(np.abs(q32 - z["q32"])) < 2e-7
    # per-quaternion norms after the column normalisation are NOT 1
    assert 0.2 < np.linalg.norm(q64, axis=2).mean() < 0.7


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_fp64_vs_reference_fp64(name):
    meta, z = load_golden(name)
    cfg = case_cfg(meta)
    params, poses = case_inputs(meta)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    d, g = onp.forward_grad(p64, poses.astype(np.float64), cfg)
    assert np.max(rel_err(d, z["d64"])) < 1e-12
    assert np.max(np.abs(g - z["g64"])) < 1e-13 * max(1.0, np.abs(z["g64"]).max() * 1e3)
    xp, dl = onp.project(p64, poses.astype(np.float64), cfg, steps=10)
    assert np.max(np.abs(xp - z["proj64"])) < 1e-12
    assert np.max(rel_err(dl, z["proj_d64"][-1])) < 1e-11


@pytest.mark.parametrize("name", CASES)
def test_numpy_oracle_fp32_within_parity_bar(name):
    """fp32 restatement vs the fp64 reference: d within 1e-5 relative (north-star bar), projected poses too."""
    meta, z = load_golden(name)
    cfg = case_cfg(meta)
    params, poses = case_inputs(meta)
    d, g = onp.forward_grad(params, poses, cfg)
    assert d.dtype == np.float32
    assert np.max(rel_err(d, z["d64"])) < 1e-5
    assert_grad_parity(g, z["g64"])
    assert_grad_parity(z["g32"], z["g64"])      # the reference's own fp32 run obeys the same bar
    xp, _ = onp.project(params, poses, cfg, steps=10)
    assert_pose_parity(xp, z["proj64"])
    assert_pose_parity(z["proj32"], z["proj64"])


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_vs_reference(name):
    meta, z = load_golden(name)
    cfg = case_cfg(meta)
    params, poses = case_inputs(meta)
    tp = otorch.to_torch_params(params, torch.float64)
    d, g, _ = otorch.forward_grad(tp, torch.from_numpy(poses).double(), cfg)
    assert np.max(rel_err(d.detach().numpy(), z["d64"])) < 1e-12
    assert np.max(np.abs(g.numpy() - z["g64"])) < 1e-12
    tp32 = otorch.to_torch_params(params, torch.float32)
    xp, _ = otorch.project(tp32, torch.from_numpy(poses), cfg, steps=10)
    # same ATen kernels as the reference in fp32 -> (near) bit-identical trajectory
    assert np.max(np.abs(xp.numpy() - z["proj32"])) < 5e-6


@pytest.mark.parametrize("name", [c for c in CASES if "noenc" not in c])
def test_train_losses_and_param_grads(name):
    meta, z = load_golden(name)
    cfg = case_cfg(meta)
    params, _ = case_inputs(meta)
    s = meta["seed"]
    tp_pose = synth.make_poses(2000 + s, 32, kind="noisy", sigma=0.25)
    tm = synth.make_poses(3000 + s, 32, kind="randn")
    tgt = (synth.uniform01(4000 + s, 32) * 0.5).astype(np.float32)
    lt = meta.get("loss_type", "l1")
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    _, ld = onp.train_losses(p64, tp_pose.astype(np.float64), tgt, tm.astype(np.float64), cfg, loss_type=lt)
    for k in ("dist", "man_loss", "eikonal"):
        assert abs(ld[k] - float(z[f"train_{k}64"])) < 1e-12 * max(1.0, abs(float(z[f"train_{k}64"])))
    tp = otorch.to_torch_params(params, torch.float64, requires_grad=True)
    _, _, grads = otorch.train_step_grads(tp, torch.from_numpy(tp_pose).double(), torch.from_numpy(tgt).double(),
                                          torch.from_numpy(tm).double(), cfg, loss_type=lt)
    names = [n for n, _ in synth.param_shapes(126, use_enc=True)]
    norms = np.array([grads[n].norm().item() for n in names])
    assert np.allclose(norms, z["train_gradnorms64"], rtol=1e-9, atol=1e-14)
    for key in z.files:
        if key.startswith("train_grad64::"):
            n = key.split("::")[1]
            assert np.allclose(grads[n].numpy(), z[key], rtol=1e-9, atol=1e-14)


def test_gradient_matches_finite_differences_fp64():
    cfg = onp.default_cfg(enc_act="softplus", df_act="softplus", enc_beta=20.0, df_beta=20.0)
    params = {k: v.astype(np.float64) for k, v in synth.make_params(11).items()}
    x = synth.make_poses(5, 3, kind="raw", dtype=np.float64)
    d, g = onp.forward_grad(params, x, cfg)
    rng = np.random.default_rng(0)
    for _ in range(6):
        v = rng.standard_normal(x.shape)
        h = 1e-6
        fd = (onp.forward(params, x + h * v, cfg) - onp.forward(params, x - h * v, cfg)) / (2 * h)
        an = np.sum(g * v, axis=(1, 2))
        assert np.allclose(fd[:, 0], an, rtol=1e-5, atol=1e-10)


def test_axis_angle_quaternion_and_vjp_fp64():
    aa = synth.make_axis_angle(3, 4, dtype=np.float64)
    aa[0, 0] = 0.0      # exercise the small-angle branch
    q = onp.axis_angle_to_quaternion(aa)
    assert np.allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-12)
    # matches torch autograd of the same formula away from 0
    t = torch.from_numpy(aa[1:]).clone().requires_grad_(True)
    ang = t.norm(dim=-1, keepdim=True)
    qt = torch.cat([torch.cos(ang / 2), t * torch.sin(ang / 2) / ang], dim=-1)
    assert np.allclose(qt.detach().numpy(), q[1:], atol=1e-13)
    w = torch.from_numpy(np.random.default_rng(1).standard_normal(qt.shape))
    (gt,) = torch.autograd.grad((qt * w).sum(), t)
    assert np.allclose(onp.axis_angle_to_quaternion_vjp(aa[1:], w.numpy()), gt.numpy(), atol=1e-12)
    # analytic limit at 0: dq/daa = [0; I/2]
    w0 = np.zeros((1, 1, 4)); w0[..., 1] = 1.0
    assert np.allclose(onp.axis_angle_to_quaternion_vjp(np.zeros((1, 1, 3)), w0), [[[0.5, 0, 0]]])


def test_batch_permutation_and_nonnegativity():
    cfg = onp.default_cfg()
    params = synth.make_params(1)
    x = synth.make_poses(9, 40)
    d = onp.forward(params, x, cfg)
    perm = np.random.default_rng(0).permutation(40)
    assert np.array_equal(onp.forward(params, x[perm], cfg), d[perm]) or np.allclose(onp.forward(params, x[perm], cfg), d[perm], rtol=1e-6)
    assert (d >= 0).all()


def test_denoise_prior_oracle_matches_torch_adam_fp64():
    """the numpy Adam loop of the oracle against torch.optim.Adam + autograd on the same prior loss."""
    params = {k: v.astype(np.float64) for k, v in synth.make_params(1).items()}
    cfg = onp.default_cfg()
    aa = synth.make_axis_angle(5, 12, dtype=np.float64).reshape(2, 6, 21, 3)
    x, d, h = onp.denoise_prior(params, aa, cfg, iterations=2, steps_per_iter=3)
    tp = otorch.to_torch_params(params, torch.float64)
    for s in range(2):
        p = torch.from_numpy(aa[s].copy()).requires_grad_(True)
        opt = torch.optim.Adam([p], 0.02, betas=(0.9, 0.999))
        for it in range(2):
            for _ in range(3):
                opt.zero_grad()
                ang = p.norm(dim=-1, keepdim=True)
                q = torch.cat([torch.cos(ang / 2), p * torch.sin(ang / 2) / ang], dim=-1)
                c = otorch.forward(tp, q, cfg).mean()
                (1e7 * c * c / (1 + it)).backward()
                opt.step()
        assert np.max(np.abs(p.detach().numpy() - x[s])) < 1e-12
    assert h.shape == (6, 2) and np.abs(x - aa).max() > 0.05
