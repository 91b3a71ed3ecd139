"""Config C5 (per-GPU shard): one trainer step (model/train_posendf.py:93-99) on 32 768 + 32 768 poses, fused path vs
plain torch autograd over the same parameters (cuBLAS), both on the GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from posendf_b200 import PoseNDF, synth

act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
def opt(fused):
    return {"train": {"device": "cuda", "loss_type": "l1", "batch_size": 4, "fused_train": fused},
            "model": {"StrEnc": {"use": True, "act": act, "beta": 100}, "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": act, "beta": 100}}}
params = {k: torch.from_numpy(v) for k, v in synth.make_params(1).items()}
tp = torch.from_numpy(synth.make_poses(1, B, kind="noisy", sigma=0.25)).cuda()
tm = torch.from_numpy(synth.make_poses(2, B)).cuda()
tgt = torch.from_numpy((synth.uniform01(3, B) * 0.5).astype(np.float32)).cuda()
out = {"act": act, "poses": B, "manifold_poses": B}
for fused in (True, False):
    net = PoseNDF(opt(fused)); net.load_state_dict(params)
    optim = torch.optim.Adam(net.parameters(), lr=1e-5, weight_decay=1e-4)
    def step():
        optim.zero_grad()
        _, ld = net(tp, tgt, tm, train=True, eikonal=1.0)
        sum(ld.values()).backward()
        optim.step()
        return ld
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); n = 10
    e0.record()
    for _ in range(n): ld = step()
    e1.record()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    out["fused" if fused else "torch_autograd"] = {"ms_per_step": dt * 1e3, "ms_per_step_device": e0.elapsed_time(e1) / n,
                                                     "samples_per_s": B / dt, "losses": {k: float(v) for k, v in ld.items()},
                                                     "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    torch.cuda.reset_peak_memory_stats()
print(json.dumps(out))
