"""Throughput of the distance-label rerank kernel (HBM-bound): queries/s and algorithmic GB/s vs the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200.engine import knn_rerank
NDB, Q, K = 2_000_000, 200_000, 500
db = torch.nn.functional.normalize(torch.randn(NDB, 21, 4, device="cuda"), dim=2)
qr = torch.nn.functional.normalize(torch.randn(Q, 21, 4, device="cuda"), dim=2)
idx = torch.randint(0, NDB, (Q, K), device="cuda", dtype=torch.int32)
knn_rerank(qr, db, idx); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); knn_rerank(qr, db, idx); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
bytes_ = Q * K * (336 + 4) + Q * (336 + 40)
peak = 6569.6
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
print(json.dumps({"kernel": "knn_rerank_kernel", "queries": Q, "candidates": K, "database_MB": NDB * 336 / 1e6, "ms": ms,
                  "queries_per_s": Q / ms * 1e3, "algorithmic_GBs": bytes_ / ms / 1e6, "hbm_peak_GBs": peak, "frac": bytes_ / ms / 1e6 / peak}))
