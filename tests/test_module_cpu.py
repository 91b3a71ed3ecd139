"""CPU: the host-side mirror of the reference interface (no compute: that needs the GPU)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from posendf_b200 import PoseNDF, synth


def make_opt(use_enc=True, act="lrelu", device="cpu", loss="l1"):
    return {"train": {"device": device, "loss_type": loss, "batch_size": 4},
            "model": {"StrEnc": {"use": use_enc, "act": act, "beta": 100},
                      "DFNet": {"in_dim": 126 if use_enc else 84, "dims": [256, 512, 1024, 512, 256, 64], "act": act, "beta": 100}}}


@pytest.mark.parametrize("use_enc", [True, False])
def test_state_dict_keys_and_shapes_match_reference(use_enc):
    ref = json.load(open(os.path.join(GOLDEN_DIR, "statedict_keys.json")))["enc" if use_enc else "noenc"]
    net = PoseNDF(make_opt(use_enc))
    mine = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    assert mine == ref
    assert [k for k, _ in mine] == [n for n, _ in synth.param_shapes(126 if use_enc else 84, use_enc=use_enc)]


def test_load_reference_style_checkpoint_and_roundtrip(tmp_path):
    params = synth.make_params(3)
    net = PoseNDF(make_opt())
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    res = net.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the checkpoint container of model/train_posendf.py:147-156
    path = tmp_path / "checkpoint_epoch_best.tar"
    torch.save({"epoch": 1, "model_state_dict": net.state_dict(), "optimizer_state_dict": {}}, path, _use_new_zipfile_serialization=False)
    net2 = PoseNDF(make_opt())
    net2.load_state_dict(torch.load(path, map_location="cpu")["model_state_dict"])
    flat = torch.cat([p.detach().reshape(-1) for p in net2._ordered_params()]).numpy()
    assert np.array_equal(flat, synth.flatten_params(params))
    assert sum(p.numel() for p in net.parameters()) == 1365565


def test_reference_call_surface():
    net = PoseNDF(make_opt())
    assert net.eval() is net or net.eval() is None      # statement form works
    net.train()
    assert hasattr(net, "enc") and hasattr(net, "dfnet") and net.device == "cpu"
    assert net.loss == "l1" and isinstance(net.loss_l1, torch.nn.L1Loss)
    assert isinstance(PoseNDF(make_opt(loss="l2")).loss_l1, torch.nn.MSELoss)
    assert PoseNDF(make_opt(use_enc=False)).enc is None


def test_eval_forward_on_cpu_fails_loudly():
    net = PoseNDF(make_opt())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(2, 21, 4), train=False)


def test_train_forward_on_cpu_fails_loudly():
    net = PoseNDF(make_opt())
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.zeros(4, 21, 4), torch.zeros(4), torch.zeros(4, 21, 4), train=True, eikonal=1.0)


def test_train_forward_matches_reference_losses():
    """the torch-autograd cross-check (tests/autograd_crosscheck.py) that the GPU tests compare the native train step with:
    in fp64 on the CPU its values must equal the reference's."""
    from autograd_crosscheck import train_forward_autograd
    from conftest import load_golden
    meta, z = load_golden("lrelu_enc_s1")
    opt = make_opt()
    net = PoseNDF(opt).double()
    net.load_state_dict({k: torch.from_numpy(v).double() for k, v in synth.make_params(1).items()})
    s = meta["seed"]
    tp = torch.from_numpy(synth.make_poses(2000 + s, 32, kind="noisy", sigma=0.25)).double()
    tm = torch.from_numpy(synth.make_poses(3000 + s, 32, kind="randn")).double()
    tgt = torch.from_numpy((synth.uniform01(4000 + s, 32) * 0.5).astype(np.float32)).double()
    loss, ld = train_forward_autograd(net, tp, tgt, tm, 1.0)
    for k in ("dist", "man_loss", "eikonal"):
        assert abs(ld[k].item() - float(z[f"train_{k}64"])) < 1e-12
    sum(ld.values()).backward()
    names = [n for n, _ in synth.param_shapes(126, use_enc=True)]
    norms = np.array([dict(net.named_parameters())[n].grad.norm().item() for n in names])
    assert np.allclose(norms, z["train_gradnorms64"], rtol=1e-9, atol=1e-14)
    _, ld0 = train_forward_autograd(net, tp, tgt, tm, 0.0)
    assert set(ld0) == {"dist"}


def test_compat_shim_resolves_reference_imports():
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from model.posendf import PoseNDF, gradient; from model.network.net_modules import DFNet, StructureEncoder, BoneMLP;"
            "from model.network.net_utils import get_parent_mapping; from configs.config import load_config;"
            "import posendf_b200; assert PoseNDF is posendf_b200.PoseNDF;"
            "assert get_parent_mapping('smpl') == [-1,-1,-1,1,2,3,4,5,6,7,8,9,9,9,12,13,14,16,17,18,19]; print('ok')"
            % (ROOT, os.path.join(ROOT, "compat")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
