"""Export the inputs of tools/tc_chain_probe.cu: the DFNet weights of a synthetic amass.yaml network and the encoder output z0
of 128 poses (fp64 oracle, stored fp32), as one flat little-endian fp32 file:
    for l in 0..6: W_l (out x in, in padded to a multiple of 32), b_l (out)      then z0 (128 x 128)
Usage: python tools/tc_chain_export.py <act: lrelu|softplus> <out.bin>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import posendf_numpy as onp
from posendf_b200 import synth

act, out = sys.argv[1], sys.argv[2]
seed = 1 if act == "lrelu" else 3
params = synth.make_params(seed)
cfg = onp.default_cfg(enc_act=act, df_act=act)
poses = synth.make_poses(1000 + seed, 128).astype(np.float64)
p64 = {k: v.astype(np.float64) for k, v in params.items()}
q, _ = onp.normalise_columns(poses)
z0, _ = onp.encoder_forward(p64, q, cfg)
widths = [126, 256, 512, 1024, 512, 256, 64, 1]
with open(out, "wb") as f:
    for l in range(7):
        W = params[f"dfnet.lin{l}.weight"].astype(np.float32)
        kp = (widths[l] + 31) // 32 * 32
        Wp = np.zeros((widths[l + 1], kp), dtype=np.float32)
        Wp[:, :widths[l]] = W
        Wp.tofile(f)
        params[f"dfnet.lin{l}.bias"].astype(np.float32).tofile(f)
    z = np.zeros((128, 128), dtype=np.float32)
    z[:, :126] = z0.astype(np.float32)
    z.tofile(f)
print("wrote", out, os.path.getsize(out), "bytes")
