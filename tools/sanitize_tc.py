"""compute-sanitizer driver for the tensor-core engine (tile policy 128 forced): forward, forward + gradient, 2-step projection, the
axis-angle prior and a short denoise loop on ragged batches, lrelu and softplus.  tools/sanitize_tc.sh runs it under memcheck, racecheck and synccheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PNDF_TILE"] = "128"
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine
for act in ("lrelu", "softplus"):
    eng = Engine(device=0, enc_act=act, df_act=act)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
    B = 128 * 3 + 37
    x = torch.from_numpy(synth.make_poses(3, B)).cuda().contiguous()
    d = eng.forward(x); d2, g = eng.forward_grad(x)
    y = x.clone(); eng.project_(y, steps=2)
    aa = torch.from_numpy(synth.make_axis_angle(4, B)).cuda()
    d3, ga = eng.prior_grad(aa)                                   # prior mode: aa -> quat prologue, VJP epilogue
    seq = torch.from_numpy(synth.make_axis_angle(5, 3 * 50)).cuda().reshape(3, 50, 21, 3).contiguous()
    d4, hist = eng.denoise_prior_(seq, iterations=1, steps_per_iter=3, lr=0.02, want_loss=True)      # fused Adam prologue, graph replay
    torch.cuda.synchronize()
    print(act, float(d.mean()), float((d - d2).abs().max()), float(g.abs().mean()), float((y - x).abs().max()), float(ga.abs().mean()),
          float(hist.mean()), eng.launch_count())
print("done")
