"""Golden vectors for the training-data feed from the REAL reference loader.

/root/reference/model/load_data.py imports pytorch3d and data.data_splits at module level (not importable here), so
`quat_flip` (load_data.py:12-16) and the class `PoseData` (load_data.py:18-86) are lifted out of the file with `ast` and
executed as they are; the instance is made with object.__new__ and given the attributes __init__ would set (file lists,
num_pts, flip).  np.random.randint is wrapped to RECORD the indices the reference draws, so that the fused feed kernel can
be given the same ones.  The input files are synthetic (posendf_b200.synth, regenerated from seeds by the test).
Build container only:  python tests/golden/make_data_golden.py"""
import ast
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from posendf_b200 import synth  # noqa: E402

N_FILES, N_AMASS, NUM_PTS = 4, 3, 96


def write_files(root):
    """the synthetic data set: returns (data_files, amass_files); tests/test_data_feed.py rebuilds the same arrays"""
    data, amass = [], []
    for i in range(N_FILES):
        d = os.path.join(root, f"ds{i % 2}")
        os.makedirs(d, exist_ok=True)
        n = 300 + 37 * i
        pose = synth.make_poses(10 + i, n, kind="noisy", sigma=0.3)
        pose[::3] *= -1                                           # negative real parts for the flip
        dist = np.abs(synth.normal(20 + i, n * 5)).reshape(n, 5).astype(np.float32)
        f = os.path.join(d, f"part{i}_000.npz")
        np.savez(f, pose=pose, dist=dist, nn_pose=pose[:, None])
        data.append(f)
    for i in range(N_AMASS):
        d = os.path.join(root, f"am{i}")
        os.makedirs(d, exist_ok=True)
        pose = synth.make_poses(50 + i, 250 + 11 * i)
        pose[1::4] *= -1
        f = os.path.join(d, f"seq{i}.npz")
        np.savez(f, pose=pose)
        amass.append(f)
    return data, amass


def main():
    src = open("/root/reference/model/load_data.py").read()
    ns = {"np": np, "torch": torch, "os": os, "Dataset": object}
    for node in ast.parse(src).body:
        if (isinstance(node, ast.FunctionDef) and node.name == "quat_flip") or (isinstance(node, ast.ClassDef) and node.name == "PoseData"):
            exec(compile(ast.Module([node], []), "load_data.py", "exec"), ns)
    PoseData = ns["PoseData"]
    out = {"N_FILES": N_FILES, "N_AMASS": N_AMASS, "NUM_PTS": NUM_PTS}
    with tempfile.TemporaryDirectory() as root:
        data, amass = write_files(root)
        for flip in (False, True):
            ds = object.__new__(PoseData)
            ds.data_files, ds.amass_files, ds.num_pts, ds.flip = data, amass, NUM_PTS, flip
            np.random.seed(1234 + int(flip))
            real = np.random.randint
            for idx in range(N_FILES):
                draws = []

                def rec(*a, **k):
                    v = real(*a, **k)
                    draws.append(np.array(v))
                    return v

                np.random.randint = rec
                try:
                    item = ds[idx]                                  # PoseData.__getitem__, unmodified
                finally:
                    np.random.randint = real
                rows, amass_idx, amass_rows = draws
                tag = f"f{int(flip)}_i{idx}"
                out[tag + "_rows"] = rows.astype(np.int64)
                out[tag + "_amass_idx"] = np.int64(amass_idx[0])
                out[tag + "_amass_rows"] = amass_rows.astype(np.int64)
                for k in ("pose", "dist", "man_poses"):
                    assert item[k].dtype == np.float32
                    out[tag + "_" + k] = item[k]
    np.savez_compressed(os.path.join(HERE, "posedata.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items() if k.startswith("f1_i0") or not k.startswith("f")})


if __name__ == "__main__":
    main()
