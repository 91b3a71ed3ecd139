import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posendf_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from posendf_b200 import synth
from posendf_b200.engine import Engine
B = 32768
eng = Engine(device=0, df_act="lrelu", enc_act="lrelu")
eng.set_weights_flat(synth.flatten_params(synth.make_params(1)))
x = torch.from_numpy(synth.make_poses(1, B)).cuda()
dump = torch.empty(B, 5504, device="cuda"); dist = torch.empty(B, 1, device="cuda"); grad = torch.empty(B, 21, 4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(sys.argv[1:] , "plain fwd+grad %.3f ms" % t(lambda: eng.forward_grad(x)),
      "export %.3f ms" % t(lambda: _lib.check(eng.lib.pndf_forward_grad_export(eng._h, x.data_ptr(), B, 1, dist.data_ptr(), grad.data_ptr(), dump.data_ptr(), None, st))))
