"""torch.optim.Adam (the reference drivers' optimizer, test_lipo.py:138-139) as one launch over all parameter tensors
(csrc/optim.cu).  Same update rule and the same per-parameter state names (`step`, `exp_avg`,
`exp_avg_sq`); the step count lives on the device, so `step()` can sit inside a captured CUDA graph."""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, ptr_array, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _group_state(self, group, dev):
        st = group.get("_mpnn_state")
        if st is None or st["step"].device != dev:
            st = dict(step=torch.zeros(1, dtype=torch.float32, device=dev),
                      ticket=torch.zeros(1, dtype=torch.int32, device=dev))
            group["_mpnn_state"] = st
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            gs = self._group_state(group, dev)
            grads, ms, vs = [], [], []
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("mpnn_b200.FusedAdam: contiguous fp32 CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("mpnn_b200.FusedAdam: sparse gradients are not supported")
                st = self.state[p]
                if not st:
                    st["step"] = gs["step"]                      # one device counter per group, shared by its tensors
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                grads.append(g)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            numel = (ctypes.c_longlong * len(ps))(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            # capacity (graph-capture) mode: the overflow flags of the edge lists built in this step gate the update on
            # the device -- a batch that exceeded the captured capacities must not touch the weights or the moments
            from . import graph
            guards = [c for c in graph._CAPTURED_COUNTS if c.device == dev][-8:] if graph._CAPACITY is not None else []
            check(lib.mpnn_adam_step(len(ps), ptr_array(ps), ptr_array(grads), ptr_array(ms), ptr_array(vs), numel,
                                     ptr(gs["step"]), ptr(gs["ticket"]), float(group["lr"]), float(b1), float(b2),
                                     float(group["eps"]), float(group["weight_decay"]), ptr_array(guards), len(guards),
                                     stream()), "adam_step")
        return loss
