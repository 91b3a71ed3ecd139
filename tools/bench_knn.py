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

# exact search (no candidate stage): fp32-FMA-bound, 21 joints x (4 mul/FMA for the dot + 1 FMA for w*|dot|) = 105 FMA per pair
from posendf_b200.engine import knn_exact, fp32_peak_tflops
del idx
for metric in ("geo", "euc"):
    QE, NE = 32768, 1_000_000
    knn_exact(qr[:QE], db[:NE], metric); torch.cuda.synchronize()
    a.record(); knn_exact(qr[:QE], db[:NE], metric); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    flops = 2.0 * 105 * QE * NE
    peak_tf = fp32_peak_tflops(0)
    print(json.dumps({"kernel": f"knn_exact_kernel<{metric}>", "queries": QE, "database_poses": NE, "ms": ms, "pairs_per_s": QE * NE / ms * 1e3,
                      "queries_per_s": QE / ms * 1e3, "algorithmic_TFLOPs": flops / ms / 1e9, "fp32_peak_TFLOPs_scalar_ffma": peak_tf,
                      "frac": flops / ms / 1e9 / peak_tf}))
