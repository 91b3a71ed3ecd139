"""Worker of tests/test_gpu_multirank.py (one process per GPU, torchrun, NCCL): the sharded paths against the unsharded ones.
Prints 'MULTIRANK OK' on rank 0 when every check passed on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from posendf_b200 import PoseNDF, synth
from posendf_b200.dist import DataParallelStep, NcclGather, PeerGather, make_gather, project_sharded
from posendf_b200.engine import Engine


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def check(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAILED: {what}", flush=True)

    eng = Engine(device=local)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
    n = 4096 + 32                                  # poses per rank
    full = torch.from_numpy(synth.make_poses(99, world * n)).to(dev)
    # unsharded reference on every rank: the whole batch on this GPU
    xs = full.clone()
    ds = eng.project_(xs, steps=3)
    # (1) gather fused into the kernel (peer stores + peer-memory barrier) == NCCL all-gather == unsharded, bit for bit
    for cls in (PeerGather, NcclGather):
        g = cls(n, dev)
        for rep in range(2):                       # twice: the barrier epochs advance
            g.local_view().copy_(full[rank * n:(rank + 1) * n])
            g.project_and_gather(eng, steps=3)
            torch.cuda.synchronize()
            dist.barrier()
            check(torch.equal(g.poses, xs), f"{cls.__name__} rep {rep}: gathered poses != unsharded")
            check(torch.equal(g.dist, ds), f"{cls.__name__} rep {rep}: gathered distances != unsharded")
        if cls is PeerGather:
            check(not g.timed_out(), "peer barrier timed out")
        dist.barrier()
        del g
    check(type(make_gather(n, dev)).__name__ == "PeerGather", "make_gather did not pick the peer path")
    # (2) ragged sharding helper (module API)
    opt = {"train": {"device": str(dev), "loss_type": "l1", "batch_size": 4},
           "model": {"StrEnc": {"use": True, "act": "lrelu", "beta": 100},
                     "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": "lrelu", "beta": 100}}}
    net = PoseNDF(opt)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(1).items()})
    m = world * n - 45                             # not a multiple of the tile or of the world size
    xg, dg = project_sharded(net, full[:m], steps=3)
    check(torch.equal(xg, xs[:m]) and torch.equal(dg, ds[:m]), "project_sharded != unsharded")
    # (3) data-parallel trainer step: identical parameters on all ranks, == the single-process step on the global batch
    B = 512
    tp = torch.from_numpy(synth.make_poses(11, world * B, kind="noisy", sigma=0.25)).to(dev)
    tm = torch.from_numpy(synth.make_poses(211, world * B)).to(dev)
    tgt = torch.from_numpy((synth.uniform01(411, world * B) * 0.5).astype(np.float32)).to(dev)
    sl = slice(rank * B, (rank + 1) * B)
    trainer = DataParallelStep(net, lr=1e-4, weight_decay=1e-4, weights=(1.0, 1.0, 1.0))
    for _ in range(2):
        trainer.step(tp[sl], tgt[sl], tm[sl])
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    check(torch.equal(ref, flat), "parameters differ across ranks after the data-parallel steps")
    single = PoseNDF(opt)
    single.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(1).items()})
    solo = DataParallelStep(single, lr=1e-4, weight_decay=1e-4, weights=(1.0, 1.0, 1.0), group=None)
    solo.world = 1                                 # the global batch on one GPU, no all-reduce
    for _ in range(2):
        solo.step(tp, tgt, tm)
    fs = torch.cat([p.detach().reshape(-1) for p in single.parameters()])
    check((fs - flat).abs().max().item() < 2e-6, f"data-parallel != single-process global batch ({(fs - flat).abs().max().item():.3e})")

    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0 and t.item() == 1:
        print("MULTIRANK OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
