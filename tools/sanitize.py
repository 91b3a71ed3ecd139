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
