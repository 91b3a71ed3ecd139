"""Golden vectors for the distance-label rerank from the REAL reference classes.

/root/reference/data/dist_utils.py imports smplx / pytorch3d at module level (not installed), so the two classes
`geo` and `euc` (dist_utils.py:9-50) are lifted out of the file with `ast` and executed as they are, with only
torch / numpy in their namespace.  Build container only."""
import ast, os, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from posendf_b200 import synth

src = open("/root/reference/data/dist_utils.py").read()
mod = ast.parse(src)
ns = {"torch": torch, "np": np}
for node in mod.body:
    if isinstance(node, ast.ClassDef) and node.name in ("geo", "euc"):
        exec(compile(ast.Module([node], []), "dist_utils.py", "exec"), ns)

Q, K, NDB = 37, 500, 4000
db = synth.make_poses(901, NDB)                                   # manifold database (unit quaternions)
qr = synth.make_poses(902, Q, kind="noisy", sigma=0.3)
idx = (synth.uniform01(903, Q * K).reshape(Q, K) * NDB).astype(np.int64)
out = {"Q": Q, "K": K, "NDB": NDB}
for name in ("geo", "euc"):
    for weighted in (False, True):
        calc = ns[name](Q, device="cpu", weighted=weighted)
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            calc.joint_weights = calc.joint_weights.to(dt)
            val, ind = calc.dist_calc(torch.from_numpy(qr).to(dt), torch.from_numpy(db[idx]).to(dt), K, 5)
            out[f"{name}_{int(weighted)}_val{tag}"] = val.numpy()
            out[f"{name}_{int(weighted)}_idx{tag}"] = ind.numpy()
# exact search = the same reference classes with the whole (small) database as every query's candidate list
NEX, QEX = 1500, 24
out["NEX"], out["QEX"] = NEX, QEX
full = np.broadcast_to(db[:NEX], (QEX, NEX, 21, 4))
for name in ("geo", "euc"):
    for weighted in (False, True):
        calc = ns[name](QEX, device="cpu", weighted=weighted)
        calc.joint_weights = calc.joint_weights.to(torch.float64)
        val, ind = calc.dist_calc(torch.from_numpy(qr[:QEX]).double(), torch.from_numpy(np.ascontiguousarray(full)).double(), NEX, 5)
        out[f"exact_{name}_{int(weighted)}_val64"] = val.numpy()
        out[f"exact_{name}_{int(weighted)}_idx64"] = ind.numpy()
np.savez_compressed(os.path.join(HERE, "knn_rerank.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "64" in str(k) or k in ("Q", "K")})
