"""Training step of the reference (model/posendf.py:78-99 losses, model/train_posendf.py:93-99 backward) on top of
the fused kernel.

    L = w_d * L1|L2(d(x), d_gt) + w_m * mean|d(x_man)| + w_e * mean_{b,j} (|g_{b,j}| - 1)^2 ,   g = d d / d x

What runs where
  * fused sm_100a kernel (libpndf):  d, g and -- exported per 32-pose tile -- every layer input z_l and every
    pre-activation adjoint  a_l = d d / d pre_l  (launch 1, also for the manifold batch without normalisation), and
    the forward-mode tangents  zdot_l  of all layer inputs along a given input tangent (launch 2, Eikonal term).
    That is 99.8 % of the per-sample arithmetic (the 7-layer DFNet chain, three times).
  * cuBLAS through torch.mm / torch.bmm (plain library GEMMs, fp32, explicit split-K for the small layers):  the
    batch reductions
        dW_l = sum_b  a_l[b] (x) (w_d delta_b z_l[b] + w_e zdot_l[b])   (+ second-order term for softplus)
    They run in backward(), when the upstream weights w_d / w_m / w_e of the three losses are known, so that the
    distance and the Eikonal term of the pose batch share ONE GEMM per layer (the manifold batch has its own adjoints).
  * two small one-thread-per-pose kernels for the 3 516-parameter structure encoder (0.2 % of the arithmetic):
    pndf_encoder_tangent (input tangent of launch 2) and pndf_encoder_param_grads (reverse sweep of the encoder for
    the first-order and the Eikonal objective incl. the softplus second-derivative terms).  No torch autograd anywhere.

Eikonal term.  With v = dE/dg held fixed, dE/dtheta = d/dtheta <v, g(theta)> = d/dtheta (JVP of d along v).  The
tangent network has the same linear structure as the linearised primal, so for layer l
        dE/dW_l = sum_b  a_l[b] (x) zdot_l[b]  +  pbar_l[b] (x) z_l[b] ,
where pbar_l is the second-order adjoint:  pbar_l = zbar_{l+1} * phi'(pre_l) + gbar_{l+1} * phi''(pre_l) * pdot_l,
zbar_l = W_l^T pbar_l.  phi'' == 0 for relu / lrelu (pbar vanishes, one extra launch is all it takes); for softplus
phi'' = beta phi' (1 - phi') and the pbar chain is evaluated here with seven cuBLAS GEMMs on the exported tensors.

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
# explicit split-K factors of the weight-gradient GEMMs (tools/tune_wgrad.py on a B200, K = 32 768 poses): the outputs are
# only 2 .. 32 tiles of 128x128, far fewer than 148 SMs
_SPLIT_K = {0: 64, 1: 16, 2: 8, 3: 8, 4: 16, 5: 64}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


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
                                                    self.grad.data_ptr(), self.dump.data_ptr(),
                                                    None if self.masks is None else self.masks.data_ptr(), _stream(x)))
        self.delta = None      # d loss / d dist per pose for unit upstream weight, (B,)
        self.v = None          # dE/dg, the pose tangent of the Eikonal term
        self.dump_t = None     # launch 2 export (tangents of the layer inputs)

    def cols(self, c0, n):
        """strided (B, n) view of dump columns [c0, c0+n) -- no copy, cuBLAS takes the row stride"""
        return self.dump[:self.B, c0:c0 + n]

    def tangent_launch(self, eng):
        x, B = self.x, self.B
        tan = torch.zeros((B + 31) // 32, 128, 32, device=x.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_encoder_tangent(eng._h, x.data_ptr(), self.v.data_ptr(), B, int(self.normalise), tan.data_ptr(),
                                                _stream(x)))
        self.dump_t = torch.empty(tan.shape[0] * 32, DUMP_ROWS, device=x.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_forward_tangent_export(eng._h, x.data_ptr(), B, int(self.normalise), tan.data_ptr(),
                                                       self.dump_t.data_ptr(),
                                                       None if self.masks is None else self.masks.data_ptr(), _stream(x)))


def _out_act_deriv(d, act, beta):
    """phi_out'(s) and phi_out''(s) recovered from d = phi_out(s): relu (relu / lrelu configs) or softplus(beta)."""
    if act == "softplus":
        sig = -torch.expm1(-beta * d)               # sigma(beta s) = 1 - exp(-beta d), accurate for tiny d
        return sig, beta * sig * (1.0 - sig)
    pos = (d > 0).to(d.dtype)
    return pos, None


def _encoder_param_grads(eng, x, v, normalise, up1, upt, upz):
    """(2, 3516) encoder parameter gradients: row 0 first-order objective, row 1 Eikonal objective"""
    out = torch.empty(2, ENC_FLOATS, device=x.device, dtype=torch.float32)
    ptr = lambda t: None if t is None else t.data_ptr()
    _lib.check(eng.lib.pndf_encoder_param_grads(eng._h, x.data_ptr(), ptr(v), x.shape[0], int(normalise), ptr(up1), ptr(upt),
                                                ptr(upz), out.data_ptr(), _stream(x)))
    return out


def _wgrad(view, a, r, split, scale=None):
    """view (n_out, n_in) += [scale *] a^T r  with a (B, n_out), r (B, n_in) row-strided; split-K over the poses through bmm"""
    B = a.shape[0]
    if split > 1 and B % split == 0 and B // split >= 512:
        prod = torch.bmm(a.unflatten(0, (split, B // split)).transpose(1, 2), r.unflatten(0, (split, B // split))).sum(0)
    elif scale is None:
        view.addmm_(a.t(), r)
        return
    else:
        prod = a.t() @ r
    if scale is None:
        view.add_(prod)
    else:
        view.addcmul_(prod, scale)


class _FlatGrads:
    """one flat fp32 gradient vector in the reference's parameter order with per-parameter views; GEMM results are
    accumulated straight into the views, the encoder kernel's 3 516 floats into the leading slice."""

    def __init__(self, net):
        ps = list(net.named_parameters())
        self.flat = torch.zeros(sum(p.numel() for _, p in ps), device=ps[0][1].device, dtype=torch.float32)
        self.views, off = {}, 0
        for n, p in ps:
            self.views[n] = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.has_enc = any(n.startswith("enc.") for n, _ in ps)

    def W(self, l):
        return self.views[f"dfnet.lin{l}.weight"]

    def b(self, l):
        return self.views[f"dfnet.lin{l}.bias"]


def _accumulate(out, net, eng, ex, up, w_eik):
    """out += up * d/dtheta sum_b delta_b d(x_b)  (+ w_eik * d/dtheta Eikonal term if ex carries a tangent export).
    up, w_eik are 0-dim device tensors (the upstream gradients of the losses); nothing here synchronises.
    ex.delta is either a (B,) tensor or a python float (the same weight for every pose: the manifold term, whose
    sign(d) is 1 wherever the adjoints are non-zero) -- then the exported layer inputs are used as they are."""
    cfg = net._cfg
    in_dim, act, beta = cfg["in_dim"], cfg["df_act"], cfg["df_beta"]
    B = ex.B
    eik = w_eik is not None and ex.dump_t is not None
    uniform = not torch.is_tensor(ex.delta)
    if uniform:
        assert not eik
        scale = up * ex.delta                      # 0-dim
        coef = scale.reshape(1, 1)
        R = ex.cols(0, Z_END)                      # strided view, no copy
    else:
        scale = None
        coef = (up * ex.delta).reshape(B, 1)
        # right-hand sides of all layers at once: R = coef * z (+ w_eik * zdot)
        R = ex.cols(0, Z_END) * coef
        if eik:
            R.addcmul_(ex.dump_t[:B, :Z_END], w_eik)
    for l in range(6):
        n_in = in_dim if l == 0 else Z_ROWS[l][1]
        _wgrad(out.W(l), ex.cols(*A_ROWS[l]), R[:, Z_ROWS[l][0]:Z_ROWS[l][0] + n_in], _SPLIT_K[l], scale)
    # biases of lin0..5: coef^T a_l, all layers in one GEMV over the adjoint columns (a_5 first, a_0 last)
    if uniform:
        ball = ex.cols(Z_END, A_END - Z_END).sum(0) * scale
    else:
        ball = (coef.t() @ ex.cols(Z_END, A_END - Z_END)).reshape(-1)
    for l in range(6):
        c0 = A_ROWS[l][0] - Z_END
        out.b(l).add_(ball[c0:c0 + A_ROWS[l][1]])
    gs, gss = _out_act_deriv(ex.dist, act, beta)                                  # (B,1)
    if uniform:
        out.W(6).addcmul_(gs.t() @ R[:, Z_ROWS[6][0]:Z_END], scale)
    else:
        out.W(6).add_(gs.t() @ R[:, Z_ROWS[6][0]:Z_END])
    out.b(6).add_((coef * gs).sum().reshape(1))
    g0 = ex.cols(G0_ROW, in_dim)                                                   # dd/dz0 per pose
    up1 = (coef * g0).contiguous()
    upt = upz = None
    if eik:
        upt = (w_eik * g0).contiguous()
        if act == "softplus":
            # second-order adjoint chain (phi'' != 0), scaled by w_eik from its seed on
            W = [getattr(net.dfnet, f"lin{l}").weight.detach() for l in range(7)]
            Z = [ex.cols(Z_ROWS[l][0], in_dim if l == 0 else Z_ROWS[l][1]) for l in range(7)]
            Zd = [ex.dump_t[:B, Z_ROWS[l][0]:Z_ROWS[l][0] + Z_ROWS[l][1]] for l in range(1, 7)]     # tangents of z_1..z_6
            pbar = (w_eik * gss) * (Zd[5] @ W[6].t())                              # adjoint of s, (B,1)
            out.W(6).add_(pbar.t() @ Z[6])
            out.b(6).add_(pbar.sum().reshape(1))
            zbar = pbar @ W[6]                                                     # (B,64) adjoint of z_6
            wdev = w_eik.reshape(1).contiguous()
            for l in range(5, -1, -1):
                # pbar_l = zbar_{l+1} phi' + w_eik phi''/phi' a_l pdot_l , one fused element-wise pass (csrc/pndf_train_ops.cuh)
                zbar = zbar.contiguous()
                n = A_ROWS[l][1]
                pbar = torch.empty(B, n, device=zbar.device, dtype=torch.float32)
                _lib.check(eng.lib.pndf_softplus_adjoint(zbar.device.index, Z[l + 1].data_ptr(), Zd[l].data_ptr(),
                                                         ex.cols(*A_ROWS[l]).data_ptr(), DUMP_ROWS, zbar.data_ptr(),
                                                         wdev.data_ptr(), float(beta), B, n, pbar.data_ptr(), _stream(zbar)))
                _wgrad(out.W(l), pbar, Z[l], _SPLIT_K[l])
                out.b(l).add_(pbar.sum(0))
                zbar = pbar @ W[l]
            upz = zbar.contiguous()
    if out.has_enc:
        eg = _encoder_param_grads(eng, ex.x, ex.v if eik else None, ex.normalise, up1, upt, upz)
        out.flat[:ENC_FLOATS].add_(eg[0])
        if eik:
            out.flat[:ENC_FLOATS].add_(eg[1])


class FusedTrainLosses(torch.autograd.Function):
    """(dist loss, manifold loss, Eikonal loss) of model/posendf.py:85-96.  forward() runs the fused launches and keeps
    their exports; backward() turns them into the parameter gradients for the upstream weights it is handed
    (model/train_posendf.py:95-98).  Memory: 22 KB per exported pose and launch (three launches per pose/manifold pair)."""

    @staticmethod
    def forward(ctx, net, pose, dist_gt, man_poses, loss_type, want_eik, *params):
        eng = net.engine()
        B = pose.shape[0]
        pose_ex, man_ex = [], []
        eik_sum = pose.new_zeros(())
        for c0 in range(0, B, CHUNK):
            ex = _Exports(eng, pose[c0:c0 + CHUNK], True, want_masks=want_eik)
            diff = ex.dist[:, 0] - dist_gt[c0:c0 + ex.B]
            ex.delta = torch.sign(diff) / B if loss_type == "l1" else 2.0 * diff / B
            if want_eik:
                nrm = ex.grad.norm(2, dim=-1, keepdim=True)
                eik_sum = eik_sum + ((nrm - 1) ** 2).sum()
                ex.v = ((2.0 * (nrm - 1) / float(B * 21)) * (ex.grad / nrm)).contiguous()   # dE/dg (mean over all (b,j))
                ex.tangent_launch(eng)
            ex.grad = None
            pose_ex.append(ex)
        d = torch.cat([e.dist for e in pose_ex], 0) if len(pose_ex) > 1 else pose_ex[0].dist
        diff = d[:, 0] - dist_gt
        loss_d = diff.abs().mean() if loss_type == "l1" else (diff * diff).mean()
        loss_m, eik = loss_d.new_zeros(()), loss_d.new_zeros(())
        if want_eik:     # the reference only reports / trains the manifold term together with the Eikonal term (posendf.py:94-99)
            Bm = man_poses.shape[0]
            msum = loss_d.new_zeros(())
            for c0 in range(0, Bm, CHUNK):
                ex = _Exports(eng, man_poses[c0:c0 + CHUNK], False)
                ex.delta = 1.0 / Bm            # sign(d) / Bm with d >= 0; where d == 0 the exported adjoints vanish
                ex.grad = None
                msum = msum + ex.dist.abs().sum()
                man_ex.append(ex)
            loss_m = msum / Bm
            eik = eik_sum / float(B * 21)
        ctx.net, ctx.eng, ctx.pose_ex, ctx.man_ex = net, eng, pose_ex, man_ex
        ctx.shapes = [p.shape for p in params]
        return loss_d, loss_m, eik

    @staticmethod
    def backward(ctx, gd, gm, ge):
        net, eng = ctx.net, ctx.eng
        if ctx.pose_ex is None:
            raise RuntimeError("FusedTrainLosses: backward() ran already and released the exports (22 KB per pose); "
                               "call the forward again instead of retain_graph=True")
        out = _FlatGrads(net)
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            for ex in ctx.pose_ex:
                _accumulate(out, net, eng, ex, gd, ge)
            for ex in ctx.man_ex:
                _accumulate(out, net, eng, ex, gm, None)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old_tf32
        ctx.pose_ex = ctx.man_ex = None
        grads, off = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            grads.append(out.flat[off:off + n].view(shp))
            off += n
        return (None, None, None, None, None, None, *grads)


def train_forward(net, pose, dist_gt, man_poses, eikonal):
    """PoseNDF.forward(train=True) on the fused path: returns (loss, dict) exactly like model/posendf.py:97-99."""
    dev = next(net.parameters()).device
    x = pose.to(dev).reshape(-1, 21, 4).float().contiguous()
    gt = dist_gt.to(dev).reshape(-1).float()
    man = man_poses.to(dev).reshape(-1, 21, 4).float().contiguous()
    params = list(net.parameters())
    loss_d, loss_m, eik = FusedTrainLosses.apply(net, x, gt, man, net.loss, eikonal > 0.0, *params)
    if eikonal > 0.0:
        return loss_d, {"dist": loss_d, "man_loss": loss_m, "eikonal": eik}
    return loss_d, {"dist": loss_d}
