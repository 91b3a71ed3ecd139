"""CPU: the host-side assembly of the fused train step (posendf_b200/train.py::_accumulate -- export column map, merged
right-hand sides, split-K through bmm, bias GEMV, the uniform-weight manifold path) on EMULATED exports.

The tensors the kernels would write (layer inputs, pre-activation adjoints, forward-mode tangents; DESIGN.md "Training
exports") are produced here with plain torch for the encoder-free configuration (DFNet eats the normalised pose,
in_dim 84), so no CUDA library call is involved; `_accumulate` must then reproduce the parameter gradients of the
reference's autograd (oracle/posendf_torch.py, pinned to the reference's golden gradients in tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import posendf_torch as otorch
from posendf_b200 import PoseNDF, synth, train

DIMS = [84, 256, 512, 1024, 512, 256, 64, 1]


def _opt(act):
    return {"train": {"device": "cpu", "loss_type": "l1", "batch_size": 4, "fused_train": False},
            "model": {"StrEnc": {"use": False, "act": act, "beta": 100},
                      "DFNet": {"in_dim": 84, "dims": DIMS[1:7], "act": act, "beta": 100}}}


class _FakeExports:
    """what train._Exports holds after launch 1 (+ the tangent launch), computed with torch on the CPU"""

    def __init__(self, params, x, act, normalise):
        slope = 0.0 if act == "relu" else 0.01
        B = x.shape[0]
        self.B, self.x, self.normalise = B, x, normalise
        W = [torch.from_numpy(params[f"dfnet.lin{l}.weight"]) for l in range(7)]
        b = [torch.from_numpy(params[f"dfnet.lin{l}.bias"]) for l in range(7)]
        self.W, self.slope = W, slope
        if normalise:
            self.n = x.norm(dim=1, keepdim=True).clamp_min(1e-12)            # F.normalize(pose, dim=1): per component over joints
            self.q = x / self.n
        else:
            self.n, self.q = torch.ones(B, 1, 4), x
        z = [self.q.reshape(B, 84)]
        self.dphi = []
        for l in range(6):
            pre = z[l] @ W[l].t() + b[l]
            self.dphi.append(torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, slope)))
            z.append(torch.where(pre > 0, pre, pre * slope))
        s = z[6] @ W[6].t() + b[6]
        self.dist = torch.relu(s)
        gs = (s > 0).float()
        a = [None] * 6
        a[5] = (gs @ W[6]) * self.dphi[5]
        for l in range(4, -1, -1):
            a[l] = (a[l + 1] @ W[l + 1]) * self.dphi[l]
        g0 = a[0] @ W[0]
        self.dump = torch.zeros(B, train.DUMP_ROWS)
        for l in range(7):
            self.dump[:, train.Z_ROWS[l][0]:train.Z_ROWS[l][0] + z[l].shape[1]] = z[l]
        for l in range(6):
            self.dump[:, train.A_ROWS[l][0]:train.A_ROWS[l][0] + a[l].shape[1]] = a[l]
        self.dump[:, train.G0_ROW:train.G0_ROW + 84] = g0
        self.grad = self._jnorm(g0.reshape(B, 21, 4))                          # d dist / d pose
        self.delta, self.v, self.dump_t = None, None, None

    def _jnorm(self, t):
        """(Jacobian of the column normalisation) applied to t -- symmetric, so also its transpose"""
        if not self.normalise:
            return t
        return (t - self.q * (self.q * t).sum(dim=1, keepdim=True)) / self.n

    def cols(self, c0, n):
        return self.dump[:self.B, c0:c0 + n]

    def tangent(self, v):
        self.v = v
        zd = [self._jnorm(v).reshape(self.B, 84)]
        for l in range(6):
            zd.append((zd[l] @ self.W[l].t()) * self.dphi[l])
        self.dump_t = torch.zeros(self.B, train.DUMP_ROWS)
        for l in range(7):
            self.dump_t[:, train.Z_ROWS[l][0]:train.Z_ROWS[l][0] + zd[l].shape[1]] = zd[l]


@pytest.mark.parametrize("act,split", [("lrelu", 1), ("lrelu", 2), ("relu", 4)])
def test_accumulate_reproduces_reference_autograd_from_emulated_exports(act, split, monkeypatch):
    monkeypatch.setattr(train, "_SPLIT_K", {l: split for l in range(6)})
    B = 2048
    params = synth.make_params(11, in_dim=84, use_enc=False)
    net = PoseNDF(_opt(act))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    tp = torch.from_numpy(synth.make_poses(2011, B, kind="noisy", sigma=0.25))
    tm = torch.from_numpy(synth.make_poses(3011, B, kind="randn"))
    gt = torch.from_numpy((synth.uniform01(4011, B) * 0.5).astype(np.float32))
    w = {"dist": 0.8, "man_loss": 0.6, "eikonal": 1.7}

    # what FusedTrainLosses.forward prepares ...
    ex = _FakeExports(params, tp, act, True)
    ex.delta = torch.sign(ex.dist[:, 0] - gt) / B
    nrm = ex.grad.norm(2, dim=-1, keepdim=True)
    ex.tangent(((2.0 * (nrm - 1) / float(B * 21)) * (ex.grad / nrm)).contiguous())
    em = _FakeExports(params, tm, act, False)
    em.delta = 1.0 / B
    # ... and what its backward does with the upstream weights
    out = train._FlatGrads(net)
    train._accumulate(out, net, None, ex, torch.tensor(w["dist"]), torch.tensor(w["eikonal"]))
    train._accumulate(out, net, None, em, torch.tensor(w["man_loss"]), None)

    cfg = dict(use_enc=False, enc_act=act, enc_beta=100.0, df_act=act, df_beta=100.0)
    tp64 = otorch.to_torch_params(params, torch.float64, requires_grad=True)
    _, ld, g_ref = otorch.train_step_grads(tp64, tp.double(), gt.double(), tm.double(), cfg, weights=w)
    assert abs(ld["dist"].item() - (ex.dist[:, 0] - gt).abs().mean().item()) < 1e-5
    assert abs(ld["eikonal"].item() - ((nrm - 1) ** 2).mean().item()) < 1e-5
    for n, ref in g_ref.items():
        got = out.views[n].double()
        err = (got - ref).norm().item() / max(ref.norm().item(), 1e-12)
        assert err < 2e-3, (n, err)          # fp32 emulation vs fp64 autograd, kink flips included; typically 1e-5


def test_flat_gradient_views_follow_the_reference_parameter_order():
    net = PoseNDF(_opt("lrelu"))
    out = train._FlatGrads(net)
    off = 0
    for (n, p), (name, shape) in zip(net.named_parameters(), synth.param_shapes(84, use_enc=False)):
        assert n == name and tuple(p.shape) == tuple(shape)
        assert out.views[n].data_ptr() == out.flat.data_ptr() + 4 * off
        off += p.numel()
    assert off == out.flat.numel() and not out.has_enc
