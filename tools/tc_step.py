"""N tensor-core projection steps over 65 536 poses (the workload ncu captures in tools/profile_tc.sh)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PNDF_TILE"] = "128"
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine
act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = Engine(device=0, enc_act=act, df_act=act)
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
x = torch.from_numpy(synth.make_poses(1, B)).cuda().contiguous()
for _ in range(n):
    eng.project_(x, steps=1)
torch.cuda.synchronize()
print("launches", eng.launch_count())
