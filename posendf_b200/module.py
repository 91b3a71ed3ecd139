"""Host-side mirror of the reference's call surface for the hot path: `PoseNDF(opt)`.

Same constructor dict, same attributes (.enc, .dfnet, .device), same forward signature and return values,
same 98 state_dict keys as /root/reference/model/posendf.py:30-101 -- so reference scripts
(experiments/sample_poses.py:71,88-93, experiments/motion_denoise.py:82,125-130) and checkpoints keep
working -- but `net(pose, train=False)` runs the fused sm_100a kernel through libpndf.so:

  * pose does not require grad  -> one forward-only launch;
  * pose requires grad          -> ONE launch computes dist and d(dist)/d(pose) analytically; backward()
                                   / torch.autograd.grad(..., create_graph=True) just scale the stored
                                   gradient by the upstream gradient (exact: dist[b] depends on pose[b] only).

There is no CPU / eager fallback on this path: without libpndf.so or without a CUDA device it raises.

The parameter-holding submodules (StructureEncoder / BoneMLP / DFNet) exist so that state_dict(),
load_state_dict(), parameters() and .to() behave as in the reference; the packed device copy of the weights
inside the engine is a cache that is rebuilt whenever a parameter tensor changes version.

train=True (model/posendf.py:78-99: dist + manifold + Eikonal losses and their parameter gradients, including the
Eikonal double backward) runs natively as well (posendf_b200/train.py): fused launches exporting the operands of the
weight-gradient reductions, a loss kernel, the split-K FFMA2 weight-gradient kernel, two small encoder kernels -- no torch
autograd graph, no library GEMM for relu / lrelu (a softplus DFNet still evaluates its second-order adjoint chain with
cuBLAS).  Gradients land in ONE flat buffer in the reference's parameter order; every `p.grad` is a view of it, and
posendf_b200.optim.FusedAdam updates the flat parameter buffer and the engine's packed weights in one kernel.  It needs
CUDA parameters and raises otherwise.  (The torch-autograd restatement of this forward that the tests cross-check
against lives in tests/autograd_crosscheck.py, not here.)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .synth import PARENTS

_BONE, _FEAT = 4, 6


def _activation(kind: str, beta: float) -> nn.Module:
    if kind == "relu":
        return nn.ReLU()
    if kind == "lrelu":
        return nn.LeakyReLU()
    if kind == "softplus":
        return nn.Softplus(beta=beta)
    raise ValueError(f"unknown activation {kind!r}")


class BoneMLP(nn.Module):
    """two-layer per-joint MLP; keys net.0 / net.2 (reference: model/network/net_modules.py:75-111)."""

    def __init__(self, has_parent: bool, act: str, beta: float):
        super().__init__()
        hidden = _BONE + _FEAT
        self.net = nn.Sequential(nn.Linear(hidden if has_parent else _BONE, hidden), _activation(act, beta),
                                 nn.Linear(hidden, _FEAT), _activation(act, beta))

    def forward(self, x):
        return self.net(x)


class StructureEncoder(nn.Module):
    """21 bone MLPs walked along the kinematic tree (reference: model/network/net_modules.py:114-170)."""

    def __init__(self, opt):
        super().__init__()
        self.parent_mapping = list(PARENTS)
        self.num_joints = len(self.parent_mapping)
        self.out_dim = self.num_joints * _FEAT
        self.net = nn.ModuleList(BoneMLP(p >= 0, opt["act"], opt["beta"]) for p in self.parent_mapping)

    def get_out_dim(self):
        return self.out_dim

    def forward(self, quat):
        feats = []
        for i, mlp in enumerate(self.net):
            p = self.parent_mapping[i]
            feats.append(mlp(quat[:, i] if p < 0 else torch.cat((quat[:, i], feats[p]), dim=-1)))
        return torch.cat(feats, dim=-1)


class DFNet(nn.Module):
    """distance head; keys lin{l} (reference: model/network/net_modules.py:9-72)."""

    def __init__(self, opt):
        super().__init__()
        widths = [opt["in_dim"], *opt["dims"], 1]
        self.num_layers = len(widths)
        for l in range(len(widths) - 1):
            setattr(self, f"lin{l}", nn.Linear(widths[l], widths[l + 1]))
        self.actv = _activation(opt["act"], opt["beta"])
        self.out_actv = nn.Softplus(beta=opt["beta"]) if opt["act"] == "softplus" else nn.ReLU()

    def forward(self, p):
        x = p.reshape(len(p), -1)
        last = self.num_layers - 2
        for l in range(last + 1):
            x = getattr(self, f"lin{l}")(x)
            x = self.actv(x) if l < last else self.out_actv(x)
        return x


class _FusedDistance(torch.autograd.Function):
    """dist = PoseNDF(pose) with the analytic input gradient computed in the same launch."""

    @staticmethod
    def forward(ctx, pose, engine, normalise):
        dist, grad = engine.forward_grad(pose, normalise=normalise)
        ctx.save_for_backward(grad)
        ctx.pose_shape = pose.shape
        return dist

    @staticmethod
    def backward(ctx, g_up):
        (grad,) = ctx.saved_tensors
        return (g_up.reshape(-1, 1, 1) * grad).reshape(ctx.pose_shape), None, None


class PoseNDF(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.device = opt["train"]["device"]
        m = opt["model"]
        self.enc = StructureEncoder(m["StrEnc"]).to(self.device) if m["StrEnc"]["use"] else None
        self.dfnet = DFNet(m["DFNet"]).to(self.device)
        self.loss = opt["train"]["loss_type"]
        self.batch_size = opt["train"]["batch_size"]
        if self.loss == "l1":
            self.loss_l1 = nn.L1Loss()
        elif self.loss == "l2":
            self.loss_l1 = nn.MSELoss()
        self._cfg = dict(use_enc=bool(m["StrEnc"]["use"]), enc_act=m["StrEnc"].get("act", "lrelu"),
                         enc_beta=float(m["StrEnc"].get("beta", 100.0)), df_act=m["DFNet"]["act"],
                         df_beta=float(m["DFNet"].get("beta", 100.0)), in_dim=int(m["DFNet"]["in_dim"]),
                         dims=tuple(int(d) for d in m["DFNet"]["dims"]))
        self._flat_param = None        # set by flatten_parameters_(): all parameters as views of one buffer
        self._flat_grad = None         # flat gradient buffer (reference parameter order); p.grad are views of it
        self._grad_views = None
        self._grad_fresh = True        # next backward() overwrites the flat gradient instead of accumulating
        self._engine = None
        self._engine_key = None
        self._weights_sig = None

    # the reference's train() override returns None; returning self is a harmless superset
    def train(self, mode=True):
        super().train(mode)
        return self

    # ------------------------------------------------------------------ fused path plumbing
    def _ordered_params(self):
        out = []
        if self.enc is not None:
            for mlp in self.enc.net:
                out += [mlp.net[0].weight, mlp.net[0].bias, mlp.net[2].weight, mlp.net[2].bias]
        for l in range(self.dfnet.num_layers - 1):
            lin = getattr(self.dfnet, f"lin{l}")
            out += [lin.weight, lin.bias]
        return out

    def engine(self):
        """libpndf handle on the device the parameters live on, with the current weights loaded."""
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("posendf_b200.PoseNDF: the distance field runs only as the fused CUDA kernel "
                               f"(parameters are on {dev}); there is no CPU fallback")
        key = (dev.index if dev.index is not None else torch.cuda.current_device())
        if self._engine is None or self._engine_key != key:
            from .engine import Engine
            c = self._cfg
            self._engine = Engine(device=key, use_enc=c["use_enc"], enc_act=c["enc_act"], enc_beta=c["enc_beta"],
                                  df_act=c["df_act"], df_beta=c["df_beta"], in_dim=c["in_dim"], dims=c["dims"])
            self._engine_key = key
            self._weights_sig = None
        sig = tuple((p.data_ptr(), p._version) for p in params)
        if sig != self._weights_sig:
            with torch.cuda.device(dev):
                self._engine.set_weights_device(torch.cat([p.detach().reshape(-1).float() for p in params]))
            self._weights_sig = sig
        return self._engine

    # ------------------------------------------------------------------ flat parameter / gradient storage (training)
    def flat_grad(self):
        """the flat fp32 gradient vector (reference parameter order) on the parameters' device, with one view per parameter"""
        params = self._ordered_params()
        dev = params[0].device
        n = sum(p.numel() for p in params)
        if self._flat_grad is None or self._flat_grad.device != dev or self._flat_grad.numel() != n:
            self._flat_grad = torch.empty(n, device=dev, dtype=torch.float32)
            self._grad_views, off = [], 0
            for p in params:
                self._grad_views.append(self._flat_grad[off:off + p.numel()].view(p.shape))
                off += p.numel()
            self._grad_fresh = True
        return self._flat_grad

    def grads_attached(self):
        """True if every p.grad is the matching view of the flat gradient buffer (then backward() accumulates into it)"""
        if self._flat_grad is None:
            return False
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self._ordered_params(), self._grad_views))

    def attach_grads(self):
        self.flat_grad()
        for p, v in zip(self._ordered_params(), self._grad_views):
            p.grad = v

    def flatten_parameters_(self):
        """Re-seat every parameter as a view of ONE flat fp32 buffer (reference order), so that the fused optimizer kernel
        and the data-parallel all-reduce work on a single contiguous vector.  state_dict / load_state_dict keep working
        (they copy in place); .to() / .double() afterwards un-flatten (call this again)."""
        params = self._ordered_params()
        if self._flat_param is not None:
            off, ok = 0, True
            for p in params:
                ok = ok and p.data_ptr() == self._flat_param.data_ptr() + 4 * off and p.dtype == torch.float32
                off += p.numel()
            if ok:
                return self._flat_param
        dev = params[0].device
        flat = torch.cat([p.detach().reshape(-1).float() for p in params]).to(dev)
        off = 0
        for p in params:
            p.data = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._flat_param = flat
        self._weights_sig = None
        return flat

    def invalidate(self):
        """Force a repack of the engine's weight copy on the next call.  The cache key is (data_ptr, _version) of every
        parameter; writes through `p.data` (p.data.copy_, EMA / clipping code) do not bump _version -- call this after
        them.  load_state_dict and optimizer steps bump the version and need nothing."""
        self._weights_sig = None

    def distance(self, pose, normalise=True):
        """(B,1) distances on the module's device; differentiable w.r.t. `pose` (first order)."""
        eng = self.engine()
        x = pose.to(device=eng.device).reshape(-1, 21, 4)
        if x.dtype != torch.float32:
            x = x.float()
        if torch.is_grad_enabled() and x.requires_grad:
            return _FusedDistance.apply(x, eng, normalise)
        return eng.forward(x, normalise=normalise)

    @torch.no_grad()
    def project(self, pose, steps=10, renorm=False):
        """the loop of experiments/sample_poses.py:70-74 fused in one launch; returns (projected poses, dist of the
        last evaluated step).  renorm=True re-normalises every quaternion after each step (north-star option)."""
        eng = self.engine()
        x = pose.detach().to(device=eng.device, dtype=torch.float32).reshape(-1, 21, 4).contiguous().clone()
        dist = eng.project_(x, steps=steps, renorm=renorm)
        return x, dist

    @torch.no_grad()
    def project_host(self, pose_cpu, steps=10, renorm=False):
        """same, host tensors in / out through pndf_project_host (copies inside the library)."""
        eng = self.engine()
        x = pose_cpu.detach().to(dtype=torch.float32).reshape(-1, 21, 4).contiguous()
        return eng.project_host(x, steps=steps, renorm=renorm)

    def prior_grad(self, axis_angle, g_up=None):
        """motion-denoise prior term (experiments/motion_denoise.py:81-83): dist(B,1) and g_up*ddist/d(axis-angle)."""
        eng = self.engine()
        return eng.prior_grad(axis_angle.to(device=eng.device), g_up=g_up)

    @torch.no_grad()
    def denoise_prior(self, axis_angle, iterations=10, steps_per_iter=50, lr=0.02, want_loss=False):
        """MotionDenoise.optimize restricted to the pose-prior term (experiments/motion_denoise.py:70-99): Adam on the
        axis-angle poses of S sequences x T frames, loss_s = 1e7/(1+it) * mean_t(dist)^2.  Returns (poses, dist, losses)."""
        eng = self.engine()
        a = axis_angle.detach().to(device=eng.device, dtype=torch.float32)
        if a.dim() == 2:
            a = a.unsqueeze(0)
        a = a.reshape(a.shape[0], a.shape[1], 63).contiguous().clone()
        dist, hist = eng.denoise_prior_(a, iterations, steps_per_iter, lr, want_loss)
        return a.reshape(a.shape[0], a.shape[1], 21, 3), dist, hist

    # ------------------------------------------------------------------ reference call surface
    def forward(self, pose, dist_gt=None, man_poses=None, train=True, eikonal=0.0):
        if not train:
            return {"dist_pred": self.distance(pose)}
        if not next(self.parameters()).is_cuda:
            raise RuntimeError("posendf_b200.PoseNDF: the train step runs only as fused CUDA kernels and needs CUDA parameters; "
                               "there is no CPU fallback")
        from .train import train_forward
        pose = pose.to(device=self.device).reshape(-1, 21, 4)
        pose.requires_grad = True                      # the reference does this in place (model/posendf.py:66)
        return train_forward(self, pose.detach(), dist_gt, man_poses, eikonal)


def gradient(inputs, outputs):
    """the reference's helper (model/posendf.py:18-27): d(outputs)/d(inputs) with ones as upstream gradient."""
    return torch.autograd.grad(outputs=outputs, inputs=inputs, grad_outputs=torch.ones_like(outputs),
                               create_graph=True, retain_graph=True, only_inputs=True)[0]
