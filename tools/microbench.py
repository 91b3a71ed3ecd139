"""Print the in-process FFMA peak (scalar FFMA and packed FFMA2) -- roofline denominator probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posendf_b200.engine import fp32_peak_tflops
for v, n in ((0, "FFMA"), (1, "FFMA2"), (2, "FFMA2+FFMA mix"), (3, "FFMA2 + LDS.128 operand traffic, 8 warps/SM, lane = mg*8+ng"), (4, "FFMA2 + LDS.128 operand traffic, 8 warps/SM, lane = ng*4+mg"), (5, "FFMA2, feature pair outermost"), (10, "mma.sync m16n8k8 tf32 (dense TFLOP/s)"), (11, "mma.sync m16n8k16 bf16")):
    print(f"fp32 peak {n}: {fp32_peak_tflops(0, v):.2f} TFLOP/s", flush=True)
