"""Generate tests/golden/*.npz by running the REAL reference modules (read-only /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference imports `ipdb` at module top; a one-line stub is put on sys.path.  Nothing is
copied from the reference: its classes are instantiated, given weights from posendf_b200.synth (so the
GPU box can regenerate the same weights without torch's RNG), and executed in fp32 and fp64 on CPU.
Inputs are regenerated from seeds at test time, so only outputs are stored.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from posendf_b200 import synth  # noqa: E402

REF = os.environ.get("POSENDF_REFERENCE", "/root/reference")


def import_reference():
    stub = tempfile.mkdtemp(prefix="ipdb_stub_")
    with open(os.path.join(stub, "ipdb.py"), "w") as f:
        f.write("def set_trace(*a, **k):\n    pass\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    from model.posendf import PoseNDF, gradient          # noqa
    from model.network.net_utils import get_parent_mapping  # noqa
    return PoseNDF, gradient, get_parent_mapping


def make_opt(case):
    in_dim = 126 if case["use_enc"] else 84
    return {
        "train": {"device": "cpu", "loss_type": case.get("loss_type", "l1"), "batch_size": 4},
        "model": {
            "StrEnc": {"use": case["use_enc"], "act": case["enc_act"], "beta": case["enc_beta"]},
            "DFNet": {"in_dim": in_dim, "dims": list(synth.AMASS_DIMS), "act": case["df_act"], "beta": case["df_beta"]},
        },
    }


CASES = [
    dict(name="lrelu_enc_s1", use_enc=True, enc_act="lrelu", df_act="lrelu", enc_beta=100, df_beta=100, seed=1, sensitised=True, pose_kind="randn"),
    dict(name="relu_enc_s2", use_enc=True, enc_act="relu", df_act="relu", enc_beta=100, df_beta=100, seed=2, sensitised=True, pose_kind="randn"),
    dict(name="softplus_enc_s3", use_enc=True, enc_act="softplus", df_act="softplus", enc_beta=100, df_beta=100, seed=3, sensitised=True, pose_kind="randn"),
    dict(name="softplus_b5_enc_s4", use_enc=True, enc_act="softplus", df_act="softplus", enc_beta=5, df_beta=5, seed=4, sensitised=True, pose_kind="noisy"),
    dict(name="lrelu_enc_default_s1", use_enc=True, enc_act="lrelu", df_act="lrelu", enc_beta=100, df_beta=100, seed=1, sensitised=False, pose_kind="rand"),
    dict(name="lrelu_noenc_s6", use_enc=False, enc_act="lrelu", df_act="lrelu", enc_beta=100, df_beta=100, seed=6, sensitised=True, pose_kind="randn"),
    dict(name="mixed_relu_softplus_s7", use_enc=True, enc_act="relu", df_act="softplus", enc_beta=100, df_beta=30, seed=7, sensitised=True, pose_kind="raw"),
    dict(name="lrelu_enc_l2_s8", use_enc=True, enc_act="lrelu", df_act="lrelu", enc_beta=100, df_beta=100, seed=8, sensitised=True, pose_kind="randn", loss_type="l2"),
]
B = 64
B_TRAIN = 32
PROJ_STEPS = 10
PROJ_STEPS_LONG = 50          # BASELINE.json configs[2]: 50-step projection loop
LONG_CASES = ("lrelu_enc_s1", "softplus_enc_s3")


def run_case(PoseNDF, gradient, case):
    in_dim = 126 if case["use_enc"] else 84
    params = synth.make_params(case["seed"], in_dim=in_dim, use_enc=case["use_enc"], sensitised=case["sensitised"])
    poses = synth.make_poses(1000 + case["seed"], B, kind=case["pose_kind"])
    out = {}
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        net = PoseNDF(make_opt(case))
        sd = {k: torch.from_numpy(v) for k, v in params.items()}
        missing = net.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        net.eval()
        if dt == torch.float64:
            net.double()
        x = torch.from_numpy(poses).to(dt).clone().requires_grad_(True)
        d = net(x, train=False)["dist_pred"]
        g = gradient(x, d)
        out["d" + tag] = d.detach().numpy()
        out["g" + tag] = g.detach().numpy()
        # the loop body of experiments/sample_poses.py:70-74, PROJ_STEPS times
        xp = torch.from_numpy(poses).to(dt).clone().requires_grad_(True)
        traj_d = []
        for _ in range(PROJ_STEPS):
            pred = net(xp, train=False)
            traj_d.append(pred["dist_pred"].detach().numpy().copy())
            gr = gradient(xp, pred["dist_pred"]).reshape(-1, 84)
            xp = (xp - (pred["dist_pred"] * gr).reshape(-1, 21, 4)).detach().requires_grad_(True)
        out["proj" + tag] = xp.detach().numpy()
        out["proj_d" + tag] = np.stack(traj_d)
        if case["name"] in LONG_CASES:
            # the same loop continued to K = 50 (BASELINE configs[2]); the first 10 steps are the trajectory above
            for _ in range(PROJ_STEPS, PROJ_STEPS_LONG):
                pred = net(xp, train=False)
                traj_d.append(pred["dist_pred"].detach().numpy().copy())
                gr = gradient(xp, pred["dist_pred"]).reshape(-1, 84)
                xp = (xp - (pred["dist_pred"] * gr).reshape(-1, 21, 4)).detach().requires_grad_(True)
            out["proj50_" + tag] = xp.detach().numpy()
            out["proj50_d" + tag] = np.stack(traj_d)
        if not case["use_enc"]:
            # the reference's train branch raises UnboundLocalError without the encoder
            # (model/posendf.py:81-83 only defines man_pose_in under `if self.enc`), so there is nothing to pin
            continue
        # train-mode forward + backward (model/posendf.py:62-99, model/train_posendf.py:93-98)
        tp = synth.make_poses(2000 + case["seed"], B_TRAIN, kind="noisy", sigma=0.25)
        tm = synth.make_poses(3000 + case["seed"], B_TRAIN, kind="randn")
        tgt = (synth.uniform01(4000 + case["seed"], B_TRAIN) * 0.5).astype(np.float32)
        net.zero_grad()
        loss, ld = net(torch.from_numpy(tp).to(dt), torch.from_numpy(tgt).to(dt), torch.from_numpy(tm).to(dt), train=True, eikonal=1.0)
        tot = sum(ld.values())
        tot.backward()
        for k, v in ld.items():
            out[f"train_{k}{tag}"] = v.detach().numpy()
        gsum = {n: p.grad.detach().numpy() for n, p in net.named_parameters()}
        out["train_gradnorms" + tag] = np.array([np.linalg.norm(gsum[n]) for n, _ in synth.param_shapes(in_dim, use_enc=case["use_enc"])])
        if dt == torch.float64:
            # a few full parameter gradients for the double-backward check
            for n in ("dfnet.lin6.weight", "dfnet.lin0.bias", "dfnet.lin3.bias") + (("enc.net.9.net.0.weight", "enc.net.0.net.2.bias") if case["use_enc"] else ()):
                out["train_grad64::" + n] = gsum[n]
    return out


def main():
    PoseNDF, gradient, get_parent_mapping = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    np.savez(os.path.join(HERE, "parents.npz"), parents=np.array(get_parent_mapping("smpl"), dtype=np.int32))
    # Q1 probe: column normalisation of the reference on a fixed input
    x = synth.make_poses(77, 8, kind="raw")
    np.savez(os.path.join(HERE, "normalise.npz"), q32=torch.nn.functional.normalize(torch.from_numpy(x), dim=1).numpy(),
             q64=torch.nn.functional.normalize(torch.from_numpy(x).double(), dim=1).numpy())
    for case in CASES:
        out = run_case(PoseNDF, gradient, case)
        meta = {k: v for k, v in case.items()}
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), meta=np.array(repr(meta)), **out)
        print(case["name"], "d32 mean/std", out["d32"].mean(), out["d32"].std(), "|g|", np.abs(out["g64"]).mean(),
              "fp32-vs-fp64 d rel", np.max(np.abs(out["d32"] - out["d64"]) / np.maximum(np.abs(out["d64"]), 1e-30)))


if __name__ == "__main__":
    main()
