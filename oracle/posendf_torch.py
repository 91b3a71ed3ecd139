"""ORACLE (test infrastructure, NOT product code) -- torch-CPU functional restatement of the PoseNDF
hot path, written the way the reference computes it: library Linear ops + autograd for d(dist)/d(pose).

This is the "port" that bench.py times as `cpu_baseline` / `--impl reference` on the GPU box's host
cores (the reference itself is Python under /root/reference and cannot travel to that box); it issues
the same ATen operator sequence the reference modules do (F.normalize, 21 bone MLPs with cat, 7 DFNet
Linears, torch.autograd.grad), so its cost is the reference's CPU cost.  Pinned against the real
reference through tests/golden/*.npz (tests/test_oracle.py).

Follows (paths under /root/reference):
  model/posendf.py:62-101, :18-27          forward / train losses / gradient helper
  model/network/net_modules.py:46-72,86-111,140-170
  model/network/net_utils.py:44-50
  experiments/sample_poses.py:70-74        projection step
  model/train_posendf.py:93-99             weighted loss + backward
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

PARENTS = (-1, -1, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19)


def _act(x, kind, beta):
    if kind == "relu":
        return F.relu(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.01)
    if kind == "softplus":
        return F.softplus(x, beta=beta)
    raise ValueError(kind)


def to_torch_params(params, dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(v).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def encoder(tp, q, cfg):
    feats = [None] * 21
    for i in range(21):
        par = PARENTS[i]
        u = q[:, i, :] if par < 0 else torch.cat((q[:, i, :], feats[par]), dim=-1)
        h = _act(F.linear(u, tp[f"enc.net.{i}.net.0.weight"], tp[f"enc.net.{i}.net.0.bias"]), cfg["enc_act"], cfg["enc_beta"])
        feats[i] = _act(F.linear(h, tp[f"enc.net.{i}.net.2.weight"], tp[f"enc.net.{i}.net.2.bias"]), cfg["enc_act"], cfg["enc_beta"])
    return torch.cat(feats, dim=-1)


def dfnet(tp, z, cfg):
    L = 0
    while f"dfnet.lin{L}.weight" in tp:
        L += 1
    x = z.reshape(len(z), -1)
    for l in range(L):
        x = F.linear(x, tp[f"dfnet.lin{l}.weight"], tp[f"dfnet.lin{l}.bias"])
        if l < L - 1:
            x = _act(x, cfg["df_act"], cfg["df_beta"])
    return _act(x, "softplus" if cfg["df_act"] == "softplus" else "relu", cfg["df_beta"])


def forward(tp, pose, cfg, normalise=True):
    x = pose.reshape(-1, 21, 4)
    q = F.normalize(x, dim=1) if normalise else x
    z = encoder(tp, q, cfg) if cfg["use_enc"] else q
    return dfnet(tp, z, cfg)


def forward_grad(tp, pose, cfg, create_graph=False):
    x = pose.detach().reshape(-1, 21, 4).clone().requires_grad_(True)
    d = forward(tp, x, cfg)
    (g,) = torch.autograd.grad(d, x, torch.ones_like(d), create_graph=create_graph, retain_graph=create_graph)
    return d, g, x


def project_step(tp, pose, cfg):
    """one body of experiments/sample_poses.py:71-74"""
    d, g, x = forward_grad(tp, pose, cfg)
    return (x - (d * g.reshape(-1, 84)).reshape(-1, 21, 4)).detach(), d.detach()


def project(tp, pose, cfg, steps=10):
    x = pose
    d = None
    for _ in range(steps):
        x, d = project_step(tp, x, cfg)
    return x, d


def train_losses(tp, pose, dist_gt, man_poses, cfg, loss_type="l1", eikonal=1.0):
    """model/posendf.py:62-99 with train=True."""
    x = pose.detach().reshape(-1, 21, 4).clone().requires_grad_(True)
    d = forward(tp, x, cfg)
    d_man = forward(tp, man_poses.reshape(-1, 21, 4), cfg, normalise=False)
    tgt = dist_gt.reshape(-1)
    loss = F.l1_loss(d[:, 0], tgt) if loss_type == "l1" else F.mse_loss(d[:, 0], tgt)
    loss_man = d_man.abs().mean()
    (g,) = torch.autograd.grad(d, x, torch.ones_like(d), create_graph=True, retain_graph=True)
    if eikonal > 0.0:
        eik = ((g.norm(2, dim=-1) - 1) ** 2).mean()
        return loss, {"dist": loss, "man_loss": loss_man, "eikonal": eik}
    return loss, {"dist": loss}


def train_step_grads(tp, pose, dist_gt, man_poses, cfg, weights=None, loss_type="l1", eikonal=1.0):
    """model/train_posendf.py:93-98: sum_k w_k * loss_k, backward -> parameter gradients."""
    weights = weights or {"dist": 1.0, "man_loss": 1.0, "eikonal": 1.0}
    for t in tp.values():
        t.grad = None
    _, ld = train_losses(tp, pose, dist_gt, man_poses, cfg, loss_type, eikonal)
    tot = sum(weights[k] * v for k, v in ld.items())
    tot.backward()
    return tot.detach(), {k: v.detach() for k, v in ld.items()}, {k: t.grad for k, t in tp.items()}
