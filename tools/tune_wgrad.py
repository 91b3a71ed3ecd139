"""Weight-gradient reductions dW_l = A_l^T R_l of the fused train step (posendf_b200/train.py): cuBLAS fp32 on the strided
export views, plain mm vs explicit split-K through bmm.  Prints ms per variant and layer; the table in train.py
(_SPLIT_K) is the argmin of this on a B200."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
dump = torch.randn(B, 5504, device=dev)
R = torch.randn(B, 2752, device=dev)
Z = [(0, 126), (128, 256), (384, 512), (896, 1024), (1920, 512), (2432, 256)]
A = [(5120, 256), (4608, 512), (3584, 1024), (3072, 512), (2816, 256), (2752, 64)]


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
for l in range(6):
    a = dump[:, A[l][0]:A[l][0] + A[l][1]]
    r = R[:, Z[l][0]:Z[l][0] + Z[l][1]]
    acc = torch.zeros(A[l][1], Z[l][1], device=dev)
    res = {"mm": timeit(lambda: acc.addmm_(a.t(), r))}
    for S in (2, 4, 8, 16, 32, 64):
        a3 = a.unflatten(0, (S, B // S)).transpose(1, 2)
        r3 = r.unflatten(0, (S, B // S))
        res[f"bmm{S}"] = timeit(lambda: acc.add_(torch.bmm(a3, r3).sum(0)))
    flops = 2.0 * B * A[l][1] * Z[l][1]
    best = min(res, key=res.get)
    out[f"layer{l}"] = {"ms": {k: round(v, 4) for k, v in res.items()}, "best": best, "best_TFLOPs": round(flops / res[best] / 1e9, 1)}
print(json.dumps(out, indent=1))
