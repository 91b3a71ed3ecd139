"""per-kernel device times of one tensor-core projection step (65 536 poses)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PNDF_TILE"] = "128"
import torch
from torch.profiler import profile, ProfilerActivity
from posendf_b200 import synth
from posendf_b200.engine import Engine
act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
eng = Engine(device=0, enc_act=act, df_act=act)
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
x = torch.from_numpy(synth.make_poses(1, B)).cuda().contiguous()
for _ in range(3):
    eng.project_(x, steps=1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.project_(x, steps=1); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
tot = 0.0
for e in evs:
    print(f"{e.device_time:9.1f} us  {e.name[:150]}")
    tot += e.device_time
print("total", tot)
