"""Latency of small batches (the reference's real call sites: B = 10 in experiments/sample_poses.py:96, one sequence of a few
hundred frames in experiments/motion_denoise.py): 32-pose tiles vs the 8-pose small-tile kernels vs the library's own choice."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine

act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
eng = Engine(device=0, enc_act=act, df_act=act)
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
out = {}
sizes = [int(v) for v in os.environ.get("PNDF_SIZES", "10,64,320,1024,1184,2368,3552,4736").split(",")]
tiles = os.environ.get("PNDF_TILES", "32,8,auto").split(",")      # add 128 for the tensor-core path
for B in sizes:
    x0 = torch.from_numpy(synth.make_poses(3, B)).cuda().contiguous()
    row = {}
    for tile in tiles:
        if tile == "auto":
            os.environ.pop("PNDF_TILE", None)
        else:
            os.environ["PNDF_TILE"] = tile
        for name, fn in (("fwd", lambda: eng.forward(x0)), ("fwd_grad_step", lambda: eng.project_(x, steps=1))):
            x = x0.clone()
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50
            a.record()
            for _ in range(n):
                fn()
            b.record(); torch.cuda.synchronize()
            row[f"{name}_tile{tile}_us"] = round(a.elapsed_time(b) / n * 1e3, 1)
    out[B] = row
    print(B, row, flush=True)
os.environ.pop("PNDF_TILE", None)
print(json.dumps(out))
