"""Fused optimizer for the trainer step (model/train_posendf.py:30,99: torch.optim.Adam(params, lr, weight_decay=1e-4)).

ONE kernel per step (adam_step_kernel, csrc/pndf_wgrad.cuh) over the module's flat parameter / gradient / moment vectors:
torch's Adam update in torch's operation order, and in the same thread the write of the new value into the engine's packed
weight buffers (slab stream: forward + reverse copy of every DFNet weight; small-parameter buffer), so the next fused launch
needs no repack.  No torch.optim, no foreach kernels, no gather kernel."""
from __future__ import annotations

import torch

from . import _lib


class FusedAdam:
    kind = "FusedAdam: one adam_step_kernel per step (moments + update + packed-weight write), flat buffers"

    def __init__(self, net, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.net = net
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.flat = net.flatten_parameters_()
        if not self.flat.is_cuda:
            raise RuntimeError("FusedAdam needs CUDA parameters (there is no CPU fallback)")
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.t = 0
        self.grad_scale = 1.0
        net.attach_grads()
        net._grad_fresh = True
        net.engine()                                   # packed copy in sync with the flat buffer before the first step

    def zero_grad(self, set_to_none=False):
        """no kernel: the next backward() overwrites the flat gradient instead of accumulating into it"""
        self.net._grad_fresh = True
        if not self.net.grads_attached():
            self.net.attach_grads()

    def flat_grad(self):
        return self.net.flat_grad()

    @torch.no_grad()
    def step(self):
        net = self.net
        if net.flatten_parameters_() is not self.flat:
            raise RuntimeError("FusedAdam: the module's parameters were re-allocated (.to() / .double()); build a new optimizer")
        eng = net.engine()                             # repacks only if somebody changed a parameter through torch
        self.t += 1
        g = net.flat_grad()
        _lib.check(eng.lib.pndf_adam_step(eng._h, self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1],
                                          self.eps, self.weight_decay, float(self.grad_scale), self.t,
                                          torch.cuda.current_stream(self.flat.device).cuda_stream))

    def state_dict(self):
        return {"step": self.t, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.t = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for k in ("lr", "eps", "weight_decay"):
            setattr(self, k, float(sd[k]))
        self.betas = tuple(float(b) for b in sd["betas"])
