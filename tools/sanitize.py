"""Small driver for compute-sanitizer runs: forward, forward+grad, 3-step projection, prior grad on a ragged batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine
for act in ("lrelu", "softplus"):
    eng = Engine(device=0, enc_act=act, df_act=act)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
    B = 32 * 5 + 7
    x = torch.from_numpy(synth.make_poses(3, B)).cuda().contiguous()
    d = eng.forward(x); d2, g = eng.forward_grad(x)
    y = x.clone(); eng.project_(y, steps=3)
    aa = torch.from_numpy(synth.make_axis_angle(4, B)).cuda()
    d3, ga = eng.prior_grad(aa)
    torch.cuda.synchronize()
    print(act, float(d.mean()), float(g.abs().mean()), float((y - x).abs().max()), float(ga.abs().mean()))
print("done")

# training step (exports, derivative handoff, tangent launch, encoder gradient kernels, softplus adjoint, device repack),
# host-buffer projection, denoise loop and rerank on ragged sizes
import numpy as np
from posendf_b200 import PoseNDF
from posendf_b200.engine import knn_rerank
for act in ("lrelu", "softplus"):
    opt = {"train": {"device": "cuda", "loss_type": "l1", "batch_size": 4},
           "model": {"StrEnc": {"use": True, "act": act, "beta": 100}, "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": act, "beta": 100}}}
    net = PoseNDF(opt)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(2).items()})
    from posendf_b200.optim import FusedAdam
    optim = FusedAdam(net, lr=1e-5, weight_decay=1e-4)
    B = 32 * 3 + 5
    tp = torch.from_numpy(synth.make_poses(5, B, kind="noisy", sigma=0.25)); tm = torch.from_numpy(synth.make_poses(6, B + 9))
    tgt = torch.from_numpy((synth.uniform01(7, B) * 0.5).astype(np.float32))
    for _ in range(2):
        optim.zero_grad()
        _, ld = net(tp, tgt, tm, train=True, eikonal=1.0)
        sum(ld.values()).backward()
        optim.step()
    xh, dh = net.project_host(torch.from_numpy(synth.make_poses(8, 1000)), steps=2)
    aa, dd, _ = net.denoise_prior(torch.from_numpy(synth.make_axis_angle(9, 3 * 11)).reshape(3, 11, 21, 3), iterations=1, steps_per_iter=2)
    torch.cuda.synchronize()
    print(act, "train", {k: float(v.detach()) for k, v in ld.items()}, float(xh.abs().mean()), float(dd.mean()))
db = torch.from_numpy(synth.make_poses(10, 500)).cuda()
q = torch.from_numpy(synth.make_poses(11, 37)).cuda()
cand = torch.randint(0, 500, (37, 50), device="cuda", dtype=torch.int32)
print("rerank", [t.shape for t in knn_rerank(q, db, cand)])
from posendf_b200.engine import knn_exact
print("exact", [t.shape for t in knn_exact(q, db, "geo")], [t.shape for t in knn_exact(q[:3], db[:133], "euc", True)])
# training-data feed kernel on ragged file sizes
import tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden.make_data_golden import write_files
from posendf_b200.data import ResidentPoseData
with tempfile.TemporaryDirectory() as root:
    data, amass = write_files(root)
    for flip in (False, True):
        feed = ResidentPoseData(data, amass, batch_size=2, num_pts=77, flip=flip, device="cuda", seed=1)
        print("feed", flip, [tuple(b["pose"].shape) for b in feed])
torch.cuda.synchronize()
print("done-train")
