"""Timings of the other BASELINE.json configs on ONE GPU (per-GPU shard sizes), not the contract bench line:
  C3: 50-step projection loop, 131 072 poses (1 048 576 / 8)          -> poses/s and pose-steps/s
  C4: motion denoise prior loop, 128 sequences x 300 frames (512x300 / 4), 100 Adam steps
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine

act = sys.argv[1] if len(sys.argv) > 1 else "softplus"
eng = Engine(device=0, enc_act=act, df_act=act)
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
out = {}
def ev():
    return torch.cuda.Event(enable_timing=True)
# ---- C3
B = 131072
x0 = torch.from_numpy(synth.make_poses(1, B)).cuda().contiguous()
x = x0.clone(); eng.project_(x, steps=2); torch.cuda.synchronize()
x = x0.clone(); a, b = ev(), ev(); a.record(); eng.project_(x, steps=50); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
out["C3_projection_50_steps"] = {"act": act, "poses_per_gpu": B, "ms": ms, "projected_poses_per_s": B / ms * 1e3, "pose_steps_per_s": B * 50 / ms * 1e3}
# ---- C4
S, T = 128, 300
aa0 = torch.from_numpy(synth.make_axis_angle(2, S * T)).cuda().reshape(S, T, 63).contiguous()
aa = aa0.clone(); eng.denoise_prior_(aa, iterations=1, steps_per_iter=2); torch.cuda.synchronize()
aa = aa0.clone(); a, b = ev(), ev(); a.record(); d, _ = eng.denoise_prior_(aa, iterations=2, steps_per_iter=50); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
out["C4_denoise_100_adam_steps"] = {"act": act, "sequences_per_gpu": S, "frames": T, "ms": ms, "sequences_per_s": S / ms * 1e3,
                                     "pose_steps_per_s": S * T * 100 / ms * 1e3, "mean_dist_before_after": None}
print(json.dumps(out))
