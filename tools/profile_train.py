"""torch-profiler view of ONE data-parallel trainer step (DataParallelStep: native losses, backward, FusedAdam) on one GPU:
which kernels run and for how long (verdict item 2: no cutlass / gemv / at::native rows for relu / lrelu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from posendf_b200 import PoseNDF, synth
from posendf_b200.dist import DataParallelStep
act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = 32768
opt = {"train": {"device": "cuda", "loss_type": "l1", "batch_size": 4},
       "model": {"StrEnc": {"use": True, "act": act, "beta": 100}, "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": act, "beta": 100}}}
net = PoseNDF(opt); net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(1).items()})
trainer = DataParallelStep(net, lr=1e-5, weight_decay=1e-4)
tp = torch.from_numpy(synth.make_poses(1, B, kind="noisy", sigma=0.25)).cuda()
tm = torch.from_numpy(synth.make_poses(2, B)).cuda()
tgt = torch.from_numpy((synth.uniform01(3, B) * 0.5).astype(np.float32)).cuda()
for _ in range(2):
    trainer.step(tp, tgt, tm)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.step(tp, tgt, tm); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
