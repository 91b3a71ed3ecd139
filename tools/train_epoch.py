"""Config C5 end to end: ONE epoch of the reference trainer loop (model/train_posendf.py:84-110 `train_model`) on
synthetic data files in the reference's training format (data/prepare_traindata.py:173: npz {'pose','dist','nn_pose'}),
fed by the device-resident loader (pndf_feed_batch), native train step, ONE all-reduce of the flat gradient vector across ranks,
FusedAdam(lr=1e-5, wd=1e-4) (posendf_b200.dist.DataParallelStep).

    python tools/train_epoch.py [--files 32] [--rows 20000] [--batch-size 8] [--num-pts 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/train_epoch.py

Every rank owns its own shard of the files (as a DistributedSampler would hand them out); one batch is
batch_size x num_pts = 32 768 poses (+ as many manifold poses) per GPU, i.e. 262 144 per step on 8 GPUs.  Rank 0 prints
one JSON line with the epoch time (max over ranks, device clock) and samples/s."""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from posendf_b200 import PoseNDF, synth
from posendf_b200.data import ResidentPoseData
from posendf_b200.dist import DataParallelStep

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=32, help="data files per rank (one item each, load_data.py:43)")
ap.add_argument("--rows", type=int, default=20000, help="poses per data file")
ap.add_argument("--batch-size", type=int, default=8)
ap.add_argument("--num-pts", type=int, default=4096)
ap.add_argument("--act", default="lrelu")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

# ---- synthetic dataset in the reference's file format (this rank's shard)
tmp = tempfile.mkdtemp(prefix=f"pndf_epoch_r{rank}_")
data_files, amass_files = [], []
for i in range(args.files):
    seed = 1000 * rank + i
    pose = synth.make_poses(seed, args.rows, kind="noisy", sigma=0.25)
    d5 = (synth.uniform01(seed + 7, args.rows * 5) * 0.5).astype(np.float32).reshape(args.rows, 5)
    f = os.path.join(tmp, f"{i:04d}000.npz")
    np.savez(f, pose=pose, dist=d5, nn_pose=np.zeros((1, 5, 21, 4), np.float32))
    data_files.append(f)
for i in range(4):
    f = os.path.join(tmp, f"amass_{i}.npz")
    np.savez(f, pose=synth.make_poses(5000 + 1000 * rank + i, args.rows))
    amass_files.append(f)
loader = ResidentPoseData(data_files, amass_files, batch_size=args.batch_size, num_pts=args.num_pts, device=dev, seed=rank)

opt = {"train": {"device": f"cuda:{local}", "loss_type": "l1", "batch_size": args.batch_size},
       "model": {"StrEnc": {"use": True, "act": args.act, "beta": 100},
                 "DFNet": {"in_dim": 126, "dims": [256, 512, 1024, 512, 256, 64], "act": args.act, "beta": 100}}}
net = PoseNDF(opt)
net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(1).items()})
net.train()
# train_posendf.py:30 Adam(lr, weight_decay=1e-4) as the fused optimizer kernel; configs/amass.yaml:56-58 loss weights 1 / 1 / 1
trainer = DataParallelStep(net, lr=1e-5, weight_decay=1e-4, weights=(1.0, 1.0, 1.0))


def epoch():
    tot, n = torch.zeros((), device=dev), 0
    for inputs in loader:                                                     # train_posendf.py:89-99, one feed-kernel launch per batch
        ld = trainer.step(inputs["pose"], inputs["dist"], inputs["man_poses"])
        tot += sum(v.detach() for v in ld.values())
        n += 1
    return tot / max(n, 1), n


epoch()                                                                       # warm-up epoch (allocator)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
avg, steps = epoch()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
per_step = args.batch_size * args.num_pts
if rank == 0:
    print(json.dumps({"config": "C5 one epoch (data feed + fused train step + grad all-reduce + Adam)", "act": args.act,
                      "n_gpus": world, "steps": steps, "samples_per_step_per_gpu": per_step, "epoch_ms": ms.item(),
                      "ms_per_step": ms.item() / steps, "samples_per_s": world * per_step * steps / ms.item() * 1e3,
                      "epoch_mean_loss": float(avg)}))
if world > 1:
    dist.destroy_process_group()
