"""Training step of the reference (model/posendf.py:78-99 losses, model/train_posendf.py:93-99 backward) on native kernels.

    L = w_d * L1|L2(d(x), d_gt) + w_m * mean|d(x_man)| + w_e * mean_{b,j} (|g_{b,j}| - 1)^2 ,   g = d d / d x

What runs where (all of it libpndf.so, include/pndf.h; this file only sequences the calls and holds the buffers)
  * forward():  per chunk of the pose batch  pndf_forward_grad_export (launch 1: d, g, and -- exported pose-major -- every
    layer input z_l and every pre-activation adjoint a_l = dd/dpre_l),  pndf_train_losses (dLoss/dd per pose, the Eikonal
    tangent v = dE/dg with torch's zero sub-gradient at |g| = 0, the loss sums),  pndf_encoder_tangent + pndf_forward_tangent_export
    (launch 2: forward-mode tangents zdot_l along v; given launch 1's activation derivatives it skips its primal pass);
    for the manifold batch launch 1 (not normalised) and the loss kernel.
  * backward(), when autograd hands over the upstream weights w_d / w_m / w_e (device scalars, nothing synchronises):
    pndf_wgrad_accumulate per chunk -- the split-K FFMA2 outer-product kernel for
        dW_l = sum_b  a_l[b] (x) (w_d delta_b z_l[b] + w_e zdot_l[b]),   db_l = sum_b w_d delta_b a_l[b]
    (one GEMM per layer and batch: the distance and the Eikonal term share it), the last layer, the encoder's reverse
    sweep, and a fixed-order reduction of the K-split partials into the module's FLAT gradient buffer; every p.grad is a
    view of that buffer (autograd's accumulation semantics are kept: fresh / None grads are overwritten, attached ones
    accumulated).
  * softplus DFNet only: phi'' != 0 adds a second-order adjoint chain
        pbar_l = zbar_{l+1} phi'(pre_l) + w_e phi''(pre_l) a_l pdot_l,   zbar_l = W_l^T pbar_l,   dW_l += pbar_l (x) z_l
    which is still evaluated with cuBLAS (torch.mm) + the fused element-wise pndf_softplus_adjoint on the exported tensors
    (phi'' == 0 for relu / lrelu -- the reference's amass.yaml -- so that configuration uses no library GEMM at all).

Eikonal term.  With v = dE/dg held fixed, dE/dtheta = d/dtheta <v, g(theta)> = d/dtheta (JVP of d along v); the tangent
network has the same linear structure as the linearised primal, which gives the zdot term above.

Everything is checked against the reference's own autograd (fp64 golden gradients) in tests/test_gpu_train.py.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

DUMP_ROWS = 5504
# export columns (pose-major dump, DESIGN.md "export column map"):
#   layer inputs z_0..z_6 at [0, 2752): (offset, width); z_0 is 126 wide with the encoder, 84 without (padded to 128)
Z_ROWS = [(0, None), (128, 256), (384, 512), (896, 1024), (1920, 512), (2432, 256), (2688, 64)]
Z_END = 2752
#   adjoints of the pre-activations pre_0..pre_5 at [2752, 5376) (pre_l feeds z_{l+1}); pre_6 is the scalar s
A_ROWS = [(5120, 256), (4608, 512), (3584, 1024), (3072, 512), (2816, 256), (2752, 64)]
A_END = 5376
G0_ROW = 5376
CHUNK = 65536      # poses per export launch (one dump buffer = 1.4 GB at this size)
ENC_FLOATS = 3516


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Exports:
    """launch 1 on one chunk of poses: distances, pose gradient and the pose-major dump of every layer input / adjoint"""

    def __init__(self, eng, x, normalise, want_masks=False):
        B = x.shape[0]
        T = (B + 31) // 32
        self.B, self.x, self.normalise = B, x, normalise
        self.dump = torch.empty(T * 32, DUMP_ROWS, device=x.device, dtype=torch.float32)
        # activation derivatives of this launch (bit masks, fp32 for softplus): the tangent launch then skips its primal pass
        self.masks = None
        if want_masks:
            n = C.c_size_t()
            _lib.check(eng.lib.pndf_act_handoff_bytes(eng._h, B, C.byref(n)))
            self.masks = torch.empty(n.value, device=x.device, dtype=torch.uint8)
        self.dist = torch.empty(B, 1, device=x.device, dtype=torch.float32)
        self.grad = torch.empty(B, 21, 4, device=x.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_forward_grad_export(eng._h, x.data_ptr(), B, int(normalise), self.dist.data_ptr(),
                                                    self.grad.data_ptr(), self.dump.data_ptr(), _ptr(self.masks), _stream(x)))
        self.coef = None       # d loss / d dist per pose for unit upstream weight, (B,); None = `uniform` for every pose
        self.uniform = 0.0
        self.v = None          # dE/dg, the pose tangent of the Eikonal term
        self.dump_t = None     # launch 2 export (tangents of the layer inputs)

    def cols(self, c0, n):
        """strided (B, n) view of dump columns [c0, c0+n) -- no copy"""
        return self.dump[:self.B, c0:c0 + n]

    def tangent_launch(self, eng):
        x, B = self.x, self.B
        tan = torch.empty((B + 31) // 32, 128, 32, device=x.device, dtype=torch.float32)
        if B % 32:
            tan[-1].zero_()                        # the kernel reads whole 32-pose tiles
        _lib.check(eng.lib.pndf_encoder_tangent(eng._h, x.data_ptr(), self.v.data_ptr(), B, int(self.normalise), tan.data_ptr(),
                                                _stream(x)))
        self.dump_t = torch.empty(tan.shape[0] * 32, DUMP_ROWS, device=x.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_forward_tangent_export(eng._h, x.data_ptr(), B, int(self.normalise), tan.data_ptr(),
                                                       self.dump_t.data_ptr(), _ptr(self.masks), _stream(x)))


def _softplus_second_order(net, eng, ex, w_eik, views):
    """Second-order adjoint chain of a softplus DFNet (phi'' != 0) for one exported chunk, scaled by w_eik from its seed on:
    adds the pbar_l (x) z_l terms into the flat-gradient views and returns upz, the second-order adjoint of z0 (B, in_dim)."""
    cfg = net._cfg
    in_dim, beta = cfg["in_dim"], cfg["df_beta"]
    B = ex.B
    W = [getattr(net.dfnet, f"lin{l}").weight.detach() for l in range(7)]
    Z = [ex.cols(Z_ROWS[l][0], in_dim if l == 0 else Z_ROWS[l][1]) for l in range(7)]
    Zd = [ex.dump_t[:B, Z_ROWS[l][0]:Z_ROWS[l][0] + Z_ROWS[l][1]] for l in range(1, 7)]     # tangents of z_1..z_6
    sig = -torch.expm1(-beta * ex.dist)                   # phi_out'(s) = sigma(beta s) = 1 - exp(-beta d)
    gss = beta * sig * (1.0 - sig)
    pbar = (w_eik * gss) * (Zd[5] @ W[6].t())             # adjoint of s, (B,1)
    views[f"dfnet.lin6.weight"].add_(pbar.t() @ Z[6])
    views[f"dfnet.lin6.bias"].add_(pbar.sum().reshape(1))
    zbar = pbar @ W[6]                                    # (B,64) adjoint of z_6
    wdev = w_eik.reshape(1).contiguous()
    for l in range(5, -1, -1):
        # pbar_l = zbar_{l+1} phi' + w_eik phi''/phi' a_l pdot_l , one fused element-wise pass (csrc/pndf_train_ops.cuh)
        zbar = zbar.contiguous()
        n = A_ROWS[l][1]
        pbar = torch.empty(B, n, device=zbar.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_softplus_adjoint(zbar.device.index, Z[l + 1].data_ptr(), Zd[l].data_ptr(),
                                                 ex.cols(*A_ROWS[l]).data_ptr(), DUMP_ROWS, zbar.data_ptr(),
                                                 wdev.data_ptr(), float(beta), B, n, pbar.data_ptr(), _stream(zbar)))
        views[f"dfnet.lin{l}.weight"].add_(pbar.t() @ Z[l])
        views[f"dfnet.lin{l}.bias"].add_(pbar.sum(0))
        zbar = pbar @ W[l]
    return zbar.contiguous()


class FusedTrainLosses(torch.autograd.Function):
    """(dist loss, manifold loss, Eikonal loss) of model/posendf.py:85-96.  forward() runs the fused launches and keeps
    their exports; backward() turns them into the parameter gradients for the upstream weights it is handed
    (model/train_posendf.py:95-98) and leaves them in the module's flat gradient buffer (p.grad = views of it).
    Memory: 22 KB per exported pose and launch (three launches per pose/manifold pair)."""

    @staticmethod
    def forward(ctx, net, pose, dist_gt, man_poses, loss_type, want_eik, *params):
        eng = net.engine()
        B = pose.shape[0]
        losses = torch.empty(3, device=pose.device, dtype=torch.float32)       # dist, Eikonal, manifold (kernel-written)
        pose_ex, man_ex = [], []
        l2 = int(loss_type != "l1")
        for c0 in range(0, B, CHUNK):
            ex = _Exports(eng, pose[c0:c0 + CHUNK], True, want_masks=want_eik)
            ex.coef = torch.empty(ex.B, device=pose.device, dtype=torch.float32)
            if want_eik:
                ex.v = torch.empty(ex.B, 21, 4, device=pose.device, dtype=torch.float32)
            _lib.check(eng.lib.pndf_train_losses(eng._h, ex.dist.data_ptr(), dist_gt[c0:c0 + ex.B].data_ptr(),
                                                 ex.grad.data_ptr() if want_eik else None, ex.B, B, 0, l2, int(c0 == 0),
                                                 ex.coef.data_ptr(), _ptr(ex.v), losses.data_ptr(), _stream(pose)))
            if want_eik:
                ex.tangent_launch(eng)
            ex.grad = None
            pose_ex.append(ex)
        if want_eik:     # the reference only reports / trains the manifold term together with the Eikonal term (posendf.py:94-99)
            Bm = man_poses.shape[0]
            for c0 in range(0, Bm, CHUNK):
                ex = _Exports(eng, man_poses[c0:c0 + CHUNK], False)
                ex.uniform = 1.0 / Bm          # sign(d) / Bm with d >= 0; where d == 0 the exported adjoints vanish
                ex.grad = None
                _lib.check(eng.lib.pndf_train_losses(eng._h, ex.dist.data_ptr(), None, None, ex.B, Bm, 1, 0, int(c0 == 0),
                                                     None, None, losses.data_ptr(), _stream(pose)))
                man_ex.append(ex)
        ctx.net, ctx.eng, ctx.pose_ex, ctx.man_ex = net, eng, pose_ex, man_ex
        loss_d, eik, loss_m = losses.unbind(0)
        return loss_d, loss_m, eik

    @staticmethod
    def backward(ctx, gd, gm, ge):
        net, eng = ctx.net, ctx.eng
        if ctx.pose_ex is None:
            raise RuntimeError("FusedTrainLosses: backward() ran already and released the exports (22 KB per pose); "
                               "call the forward again instead of retain_graph=True")
        flat = net.flat_grad()
        fresh = net._grad_fresh
        if not net.grads_attached():
            # p.grad is None (zero_grad(set_to_none=True), the torch default): overwrite.  Foreign gradient tensors (another
            # autograd path wrote them) are carried over into the flat buffer first, so accumulation semantics hold.
            params = net._ordered_params()
            fresh = all(p.grad is None for p in params)
            if not fresh:
                for p, v in zip(params, net._grad_views):
                    if p.grad is None:
                        v.zero_()
                    elif p.grad.data_ptr() != v.data_ptr():
                        v.copy_(p.grad)
        softplus_eik = net._cfg["df_act"] == "softplus" and any(ex.dump_t is not None for ex in ctx.pose_ex)
        if fresh and softplus_eik:
            flat.zero_()                 # the second-order chain below adds into the views before the first kernel call
            fresh = False
        views = dict(zip((n for n, _ in net.named_parameters()), net._grad_views)) if softplus_eik else None
        gd, gm, ge = (g.detach().to(torch.float32).contiguous() for g in (gd, gm, ge))
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            for ex in ctx.pose_ex:
                eik = ex.dump_t is not None
                upz = _softplus_second_order(net, eng, ex, ge, views) if (softplus_eik and eik) else None
                _lib.check(eng.lib.pndf_wgrad_accumulate(eng._h, ex.x.data_ptr(), _ptr(ex.v), 1, ex.dump.data_ptr(), _ptr(ex.dump_t),
                                                         ex.coef.data_ptr(), 0.0, ex.dist.data_ptr(), ex.B, gd.data_ptr(),
                                                         ge.data_ptr() if eik else None, _ptr(upz), flat.data_ptr(), int(fresh),
                                                         _stream(flat)))
                fresh = False
            for ex in ctx.man_ex:
                _lib.check(eng.lib.pndf_wgrad_accumulate(eng._h, ex.x.data_ptr(), None, 0, ex.dump.data_ptr(), None, None,
                                                         float(ex.uniform), ex.dist.data_ptr(), ex.B, gm.data_ptr(), None, None,
                                                         flat.data_ptr(), int(fresh), _stream(flat)))
                fresh = False
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old_tf32
        ctx.pose_ex = ctx.man_ex = None
        net._grad_fresh = False
        if not net.grads_attached():
            net.attach_grads()
        # the gradients are already where autograd would put them (p.grad views of the flat buffer): nothing to hand back
        return (None,) * (6 + len(net._grad_views))


def train_forward(net, pose, dist_gt, man_poses, eikonal):
    """PoseNDF.forward(train=True) on the fused path: returns (loss, dict) exactly like model/posendf.py:97-99."""
    dev = next(net.parameters()).device
    x = pose.to(dev).reshape(-1, 21, 4).float().contiguous()
    gt = dist_gt.to(dev).reshape(-1).float().contiguous()
    man = man_poses.to(dev).reshape(-1, 21, 4).float().contiguous()
    params = list(net.parameters())
    loss_d, loss_m, eik = FusedTrainLosses.apply(net, x, gt, man, net.loss, eikonal > 0.0, *params)
    if eikonal > 0.0:
        return loss_d, {"dist": loss_d, "man_loss": loss_m, "eikonal": eik}
    return loss_d, {"dist": loss_d}
