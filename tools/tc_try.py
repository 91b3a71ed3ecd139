import os, sys
sys.path.insert(0, "/root/repo")
os.environ["PNDF_TILE"] = "128"
import numpy as np, torch
from posendf_b200 import synth
from posendf_b200.engine import Engine
from oracle import posendf_numpy as onp
eng = Engine(device=0)
params = synth.make_params(1)
eng.set_weights_flat(synth.flatten_params(params))
x = torch.from_numpy(synth.make_poses(3, 300)).cuda()
torch.cuda.synchronize()
try:
    d = eng.forward(x); torch.cuda.synchronize(); print("fwd ok", d[:4].flatten().tolist())
    d2, g = eng.forward_grad(x); torch.cuda.synchronize()
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    dref, gref = onp.forward_grad(p64, synth.make_poses(3, 300).astype(np.float64), onp.default_cfg())
    print("d rel err", np.max(np.abs(d2.cpu().numpy() - dref) / np.abs(dref)))
    e = np.linalg.norm((g.cpu().numpy() - gref).reshape(300, -1), axis=1) / np.linalg.norm(gref.reshape(300, -1), axis=1)
    print("grad err median/max", np.median(e), e.max())
except Exception as ex:
    print("ERR", ex)
