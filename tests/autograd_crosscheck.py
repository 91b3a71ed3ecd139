"""TEST INFRASTRUCTURE -- torch-autograd restatement of PoseNDF.forward(train=True) over the product module's parameter
submodules (the reference's model/posendf.py:62-99 written against posendf_b200.PoseNDF's .enc / .dfnet), used by the tests
as the cross-check of the native train step: same parameters, plain torch autograd (cuBLAS on a GPU, ATen on the CPU).
It lived in posendf_b200/module.py in round 1; it is not a product path."""
import torch
import torch.nn as nn


def train_forward_autograd(net, pose, dist_gt, man_poses, eikonal):
    dev = next(net.parameters()).device
    dt = next(net.parameters()).dtype
    pose = pose.to(device=dev, dtype=dt).reshape(-1, 21, 4)
    pose.requires_grad = True
    dist_gt = dist_gt.to(device=dev, dtype=dt).reshape(-1)
    q = nn.functional.normalize(pose, dim=1)
    dist_pred = net.dfnet(net.enc(q) if net.enc is not None else q)
    man = man_poses.to(device=dev, dtype=dt).reshape(-1, 21, 4)
    dist_man = net.dfnet(net.enc(man) if net.enc is not None else man)
    loss = net.loss_l1(dist_pred[:, 0], dist_gt)
    if eikonal > 0.0:
        (g,) = torch.autograd.grad(dist_pred, pose, torch.ones_like(dist_pred), create_graph=True, retain_graph=True)
        eik = ((g.norm(2, dim=-1) - 1) ** 2).mean()
        return loss, {"dist": loss, "man_loss": dist_man.abs().mean(), "eikonal": eik}
    return loss, {"dist": loss}
