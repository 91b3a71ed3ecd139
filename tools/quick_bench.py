"""Quick timing of the fused kernel (not the contract bench): poses/s for fwd, fwd+grad+step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200 import synth
from posendf_b200.engine import Engine
act = sys.argv[1] if len(sys.argv) > 1 else "lrelu"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
eng = Engine(device=0, enc_act=act, df_act=act)
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
x = torch.from_numpy(synth.make_poses(1, B)).cuda().contiguous()
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: eng.forward(x)); print(f"{act} B={B} forward      : {t:8.3f} ms  {B/t*1e3:.3e} poses/s  {B*2.725e6/t/1e9:.1f} TFLOP/s")
y = x.clone()
t = timeit(lambda: eng.project_(y, steps=1)); print(f"{act} B={B} fwd+grad+step: {t:8.3f} ms  {B/t*1e3:.3e} poses/s  {B*5.45e6/t/1e9:.1f} TFLOP/s")
