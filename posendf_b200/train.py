"""Training step of the reference (model/posendf.py:78-99 losses, model/train_posendf.py:93-99 backward) on top of
the fused kernel.

    L = w_d * L1|L2(d(x), d_gt) + w_m * mean|d(x_man)| + w_e * mean_{b,j} (|g_{b,j}| - 1)^2 ,   g = d d / d x

What runs where
  * fused sm_100a kernel (libpndf):  d, g and -- exported per 32-pose tile -- every layer input z_l and every
    pre-activation adjoint  a_l = d d / d pre_l  (launch 1, also for the manifold batch without normalisation), and
    the forward-mode tangents  zdot_l  of all layer inputs along a given input tangent (launch 2, Eikonal term).
    That is 99.8 % of the per-sample arithmetic (the 7-layer DFNet chain, three times).
  * cuBLAS through torch.mm (plain library GEMMs, fp32):  the batch reductions
        dW_l = sum_b  a_l[b] (x) (delta_b z_l[b] + zdot_l[b])   (+ second-order term for softplus)
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
# layer inputs z_0..z_6: (row offset, width); z_0 is 126 wide with the encoder, 84 without (rows padded to 128)
Z_ROWS = [(0, None), (128, 256), (384, 512), (896, 1024), (1920, 512), (2432, 256), (2688, 64)]
# adjoints of the pre-activations pre_0..pre_5 (pre_l feeds z_{l+1}); pre_6 is the scalar s
A_ROWS = [(5120, 256), (4608, 512), (3584, 1024), (3072, 512), (2816, 256), (2752, 64)]
G0_ROW = 5376
CHUNK = 65536      # poses per export chunk (bounds the dump buffers at ~1.4 GB each)


def _rows(dump, r0, n, B):
    """dump (Bpad, 5504) pose-major -> strided (B, n) view of columns [r0, r0+n) (no copy; cuBLAS takes the stride)"""
    return dump[:B, r0:r0 + n]


class _Exports:
    def __init__(self, eng, x, normalise):
        B = x.shape[0]
        T = (B + 31) // 32
        self.B = B
        self.dump = torch.empty(T * 32, DUMP_ROWS, device=x.device, dtype=torch.float32)
        self.dist = torch.empty(B, 1, device=x.device, dtype=torch.float32)
        self.grad = torch.empty(B, 21, 4, device=x.device, dtype=torch.float32)
        _lib.check(eng.lib.pndf_forward_grad_export(eng._h, x.data_ptr(), B, int(normalise), self.dist.data_ptr(),
                                                    self.grad.data_ptr(), self.dump.data_ptr(),
                                                    torch.cuda.current_stream(x.device).cuda_stream))

    def z(self, l, in_dim):
        r0, n = Z_ROWS[l]
        return _rows(self.dump, r0, in_dim if l == 0 else n, self.B)

    def a(self, l):
        r0, n = A_ROWS[l]
        return _rows(self.dump, r0, n, self.B)

    def g0(self, in_dim):
        return _rows(self.dump, G0_ROW, in_dim, self.B)


def _out_act_deriv(d, act, beta):
    """phi_out'(s) and phi_out''(s) recovered from d = phi_out(s): relu (relu / lrelu configs) or softplus(beta)."""
    if act == "softplus":
        sig = -torch.expm1(-beta * d)               # sigma(beta s) = 1 - exp(-beta d), accurate for tiny d
        return sig, beta * sig * (1.0 - sig)
    pos = (d > 0).to(d.dtype)
    return pos, torch.zeros_like(d)


def _hidden_act_deriv(z_next, act, beta):
    """phi'(pre_l) and phi''(pre_l)/phi'(pre_l) recovered from z_{l+1} = phi(pre_l)."""
    if act == "softplus":
        sig = -torch.expm1(-beta * z_next)           # phi' = sigma(beta pre) = 1 - exp(-beta z), accurate for tiny z
        return sig, beta * (1.0 - sig)
    slope = 0.0 if act == "relu" else 0.01
    return torch.where(z_next > 0, torch.ones_like(z_next), torch.full_like(z_next, slope)), None


ENC_FLOATS = 3516


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _encoder_tangent(eng, x, v, normalise):
    """tangent of the DFNet input along the pose tangent v, already in the [tile][128][32] layout of launch 2"""
    B = x.shape[0]
    tiles = torch.zeros((B + 31) // 32, 128, 32, device=x.device, dtype=torch.float32)
    _lib.check(eng.lib.pndf_encoder_tangent(eng._h, x.data_ptr(), v.data_ptr(), B, int(normalise), tiles.data_ptr(), _stream(x)))
    return tiles


def _encoder_param_grads(eng, x, v, normalise, up1, upt, upz):
    """(2, 3516) encoder parameter gradients: row 0 first-order objective, row 1 Eikonal objective"""
    out = torch.empty(2, ENC_FLOATS, device=x.device, dtype=torch.float32)
    ptr = lambda t: None if t is None else t.data_ptr()
    _lib.check(eng.lib.pndf_encoder_param_grads(eng._h, x.data_ptr(), ptr(v), x.shape[0], int(normalise), ptr(up1), ptr(upt),
                                                ptr(upz), out.data_ptr(), _stream(x)))
    return out


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

    def addmm(self, name, a_t, b):          # view += a_t @ b
        self.views[name].addmm_(a_t, b)

    def add(self, name, val):
        v = self.views[name]
        v.add_(val.reshape(v.shape))

    def add_encoder(self, flat3516):
        self.flat[:ENC_FLOATS].add_(flat3516)


def fused_param_grads(net, x, delta_fn, normalise=True, eik_weight=None, loss_norm=None):
    """Parameter gradients of  sum_b delta[b] * d(x[b])  (+ those of the Eikonal term if eik_weight is not None), the
    distances and the Eikonal value.  x (B,21,4) fp32 CUDA; delta_fn(dist_chunk, lo, hi) -> (hi-lo,) upstream gradient
    on d for poses [lo,hi) (evaluated after the distances are known, so no extra forward launch is needed).  Returns
    (first-order _FlatGrads, Eikonal _FlatGrads or None, dist (B,1), eikonal scalar or None)."""
    eng = net.engine()
    cfg = net._cfg
    in_dim, act, beta = cfg["in_dim"], cfg["df_act"], cfg["df_beta"]
    W = [getattr(net.dfnet, f"lin{l}").weight.detach() for l in range(7)]
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g1, ge = _FlatGrads(net), (_FlatGrads(net) if eik_weight is not None else None)
        dists, eik_sum = [], x.new_zeros(())
        Btot = x.shape[0]
        for c0 in range(0, Btot, CHUNK):
            xc = x[c0:c0 + CHUNK].contiguous()
            ex = _Exports(eng, xc, normalise)
            dists.append(ex.dist)
            dl = delta_fn(ex.dist, c0, c0 + xc.shape[0]).reshape(-1, 1)
            gs, gss = _out_act_deriv(ex.dist, act, beta)                    # (B,1)
            A = [ex.a(l) for l in range(6)]
            Z = [ex.z(l, in_dim) for l in range(7)]
            # ---- first-order term: dW_l = (delta * a_l)^T z_l
            for l in range(6):
                da = dl * A[l]
                g1.addmm(f"dfnet.lin{l}.weight", da.t(), Z[l])
                g1.add(f"dfnet.lin{l}.bias", da.sum(0))
            g1.add("dfnet.lin6.weight", ((dl * gs) * Z[6]).sum(0))
            g1.add("dfnet.lin6.bias", (dl * gs).sum(0))
            g0 = ex.g0(in_dim).contiguous()                                  # adjoint of the encoder output (unit upstream)
            up0 = dl * g0
            if eik_weight is None:
                if net.enc is not None:
                    g1.add_encoder(_encoder_param_grads(eng, xc, None, normalise, up0, None, None)[0])
                continue
            # ---- Eikonal term
            g = ex.grad
            nrm = g.norm(2, dim=-1, keepdim=True)
            count = float((loss_norm if loss_norm is not None else Btot) * 21)
            eik_sum = eik_sum + ((nrm - 1) ** 2).sum()
            v = ((2.0 * (nrm - 1) / count) * (g / nrm)).contiguous()         # dE/dg  (mean over all (b,j))
            tan = _encoder_tangent(eng, xc, v, normalise)
            dump_t = torch.empty(tan.shape[0] * 32, DUMP_ROWS, device=xc.device, dtype=torch.float32)
            _lib.check(eng.lib.pndf_forward_tangent_export(eng._h, xc.data_ptr(), ex.B, int(normalise), tan.data_ptr(),
                                                           dump_t.data_ptr(), _stream(xc)))
            Zd = [_rows(dump_t, 0, in_dim, ex.B)] + [_rows(dump_t, Z_ROWS[l][0], Z_ROWS[l][1], ex.B) for l in range(1, 7)]
            for l in range(6):
                ge.addmm(f"dfnet.lin{l}.weight", A[l].t(), Zd[l])
            ge.add("dfnet.lin6.weight", (gs * Zd[6]).sum(0))
            up_z0 = None
            if act == "softplus":
                # second-order adjoint chain (phi'' != 0): seven GEMMs on the exported tensors
                sdot = Zd[6] @ W[6].t()                                      # (B,1)
                pbar = gss * sdot                                            # adjoint of s
                ge.add("dfnet.lin6.weight", (pbar * Z[6]).sum(0))
                ge.add("dfnet.lin6.bias", pbar.sum(0))
                zbar = pbar @ W[6]                                           # (B,64) adjoint of z_6
                for l in range(5, -1, -1):
                    d1, ratio = _hidden_act_deriv(Z[l + 1], act, beta)       # phi'(pre_l), phi''/phi'
                    pdot = Zd[l + 1] / d1.clamp_min(1e-30)                   # tangent of pre_l
                    pbar = zbar * d1 + A[l] * ratio * pdot                   # A[l] = gbar_{l+1} * phi'
                    ge.addmm(f"dfnet.lin{l}.weight", pbar.t(), Z[l])
                    ge.add(f"dfnet.lin{l}.bias", pbar.sum(0))
                    zbar = pbar @ W[l]
                up_z0 = zbar.contiguous()
            if net.enc is not None:
                eg = _encoder_param_grads(eng, xc, v, normalise, up0, g0, up_z0)
                g1.add_encoder(eg[0])
                ge.add_encoder(eg[1])
        dist = torch.cat(dists, 0) if len(dists) > 1 else dists[0]
        eik = None
        if eik_weight is not None:
            eik = eik_sum / float((loss_norm if loss_norm is not None else Btot) * 21)
        return g1, ge, dist, eik
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32


class FusedTrainLosses(torch.autograd.Function):
    """(dist loss, manifold loss, Eikonal loss) of model/posendf.py:85-96 with parameter gradients from the fused path.
    backward() combines the three per-term gradient vectors with the upstream weights (model/train_posendf.py:95-98)."""

    @staticmethod
    def forward(ctx, net, pose, dist_gt, man_poses, loss_type, want_eik, *params):
        B = pose.shape[0]

        def delta_dist(dist, lo, hi):
            diff = dist[:, 0] - dist_gt[lo:hi]
            return torch.sign(diff) / B if loss_type == "l1" else 2.0 * diff / B

        g_dist, g_eik, d, eik = fused_param_grads(net, pose, delta_dist, True, 1.0 if want_eik else None)
        diff = d[:, 0] - dist_gt
        loss_d = diff.abs().mean() if loss_type == "l1" else (diff * diff).mean()
        g_man, loss_m = None, loss_d.new_zeros(())
        if want_eik:     # the reference only reports / trains the manifold term together with the Eikonal term (posendf.py:94-99)
            Bm = man_poses.shape[0]
            g_man, _, dm, _ = fused_param_grads(net, man_poses, lambda dist, lo, hi: torch.sign(dist[:, 0]) / Bm, False, None)
            loss_m = dm.abs().mean()
        ctx.flats = [g.flat if g is not None else None for g in (g_dist, g_man, g_eik)]
        ctx.shapes = [p.shape for p in params]
        if not want_eik:
            eik = loss_d.new_zeros(())
        return loss_d, loss_m, eik

    @staticmethod
    def backward(ctx, gd, gm, ge):
        tot = None
        for up, flat in zip((gd, gm, ge), ctx.flats):
            if flat is None or up is None:
                continue
            tot = up * flat if tot is None else tot.add_(up * flat)
        out, off = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            out.append(tot[off:off + n].view(shp))
            off += n
        return (None, None, None, None, None, None, *out)


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
