"""Config C5: data-parallel trainer step, 32 768 + 32 768 samples per GPU, fused train step on every
rank + ONE all-reduce of the 1 365 565 parameter gradients + Adam (model/train_posendf.py:30,93-99).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_train_ddp.py [act] [samples_per_gpu]

Rank 0 prints one JSON line: whole-job samples/s (max-over-ranks device time) and a check that all ranks hold
bit-identical parameters after the timed steps."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from posendf_b200 import PoseNDF, synth
from posendf_b200.dist import allreduce_gradients

act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opt = {"train": {"device": f"cuda:{local}", "loss_type": "l1", "batch_size": 4},
       "model": {"StrEnc": {"use": True, "act": act, "beta": 100}, "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": act, "beta": 100}}}
net = PoseNDF(opt)
net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(1).items()})
optim = torch.optim.Adam(net.parameters(), lr=1e-5, weight_decay=1e-4)
dev = torch.device("cuda", local)
tp = torch.from_numpy(synth.make_poses(11 + rank, B, kind="noisy", sigma=0.25)).to(dev)
tm = torch.from_numpy(synth.make_poses(211 + rank, B)).to(dev)
tgt = torch.from_numpy((synth.uniform01(411 + rank, B) * 0.5).astype(np.float32)).to(dev)


def step():
    optim.zero_grad()
    _, ld = net(tp, tgt, tm, train=True, eikonal=1.0)
    sum(ld.values()).backward()
    if world > 1:
        allreduce_gradients(net)
    optim.step()
    return ld


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ld = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
same = True
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = flat.clone()
    dist.broadcast(ref, 0)
    eq = torch.tensor([int(torch.equal(ref, flat))], device=dev)
    dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    same = bool(eq.item())
if rank == 0:
    print(json.dumps({"config": "C5 data-parallel train step", "act": act, "n_gpus": world, "samples_per_gpu": B,
                      "ms_per_step": ms.item(), "samples_per_s": world * B / ms.item() * 1e3,
                      "params_identical_across_ranks": same, "losses_rank0": {k: float(v) for k, v in ld.items()}}))
if world > 1:
    dist.destroy_process_group()
