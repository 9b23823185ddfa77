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

    # ---- data parallel: gradient all-reduce fused into the step over NVLink peer memory (csrc/optim.cu k_adam_ddp) ----
    def enable_ddp(self, group=None):
        """After this, `step()` sums the gradients over the ranks of `group` (default: the world) INSIDE the Adam launch:
        every rank's gradients are packed into a symmetric-memory buffer that all peers map, chunks are published with
        system-scope flags and summed in rank order (one gradient sum per step, nothing else).  Replaces
        pack + ncclAllReduce + scale + unpack + Adam (5 launches, a latency-bound collective) by ONE launch.
        Call it on every rank, after the parameters are on their device and before the first step."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return False
        self._ddp_on = True
        if len(self.param_groups) != 1:
            raise RuntimeError("mpnn_b200.FusedAdam.enable_ddp: one parameter group only")
        self._ddp_group = group
        self._ddp = None      # buffers are built at the first step, from the parameters that receive gradients
        return True

    def _build_ddp(self, ps):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self._ddp_group
        lib = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8 or len(ps) > 72:
            raise RuntimeError("mpnn_b200.FusedAdam.enable_ddp: at most 8 ranks (one NVSwitch domain) and 72 tensors")
        dev = ps[0].device
        goff, o = [], 0
        for p in ps:
            goff.append(o)
            o += (p.numel() + 3) // 4 * 4
        region = o
        numel = (ctypes.c_longlong * len(ps))(*[p.numel() for p in ps])
        n_chunks = lib.mpnn_adam_step_ddp(len(ps), None, None, None, None, numel, None, None, None, 0.0, 0.0, 0.0, 0.0, 0.0,
                                          None, None, 0, world, rank, None)
        total = 2 * region + world * n_chunks + 64
        flat = symm_mem.empty((total,), dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(flat, group if group is not None else dist.group.WORLD)
        flat.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group)          # no rank may publish a flag into a buffer that is still being zeroed
        bufs = [int(b) for b in hdl.buffer_ptrs]
        self._ddp = dict(params=ps, flat=flat, hdl=hdl, world=world, rank=rank, region=region, n_chunks=n_chunks,
                         goff=(ctypes.c_longlong * len(ps))(*goff),
                         flat_ptrs=(ctypes.c_void_p * world)(*bufs),
                         flag_ptrs=(ctypes.c_void_p * world)(*[b + 2 * region * 4 for b in bufs]))

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
            if getattr(self, "_ddp_on", False):
                if self._ddp is None:
                    self._build_ddp(ps)     # collective: every rank reaches its first step
                dd = self._ddp
                if [id(p) for p in ps] != [id(p) for p in dd["params"]]:
                    raise RuntimeError("mpnn_b200.FusedAdam (ddp): every parameter needs a gradient on every rank, every step")
                check(lib.mpnn_adam_step_ddp(len(ps), ptr_array(ps), ptr_array(grads), ptr_array(ms), ptr_array(vs), numel,
                                             dd["goff"], ptr(gs["step"]), ptr(gs["ticket"]), float(group["lr"]), float(b1),
                                             float(b2), float(group["eps"]), float(group["weight_decay"]),
                                             ctypes.cast(dd["flat_ptrs"], ctypes.POINTER(ctypes.c_void_p)),
                                             ctypes.cast(dd["flag_ptrs"], ctypes.POINTER(ctypes.c_void_p)), dd["region"],
                                             dd["world"], dd["rank"], stream()), "adam_step_ddp")
                continue
            # capacity (graph-capture) mode: the overflow flags of the edge lists built in this step gate the update on
            # the device -- a batch that exceeded the captured capacities must not touch the weights or the moments
            from . import graph
            guards = [c for c in graph._CAPTURED_COUNTS if c.device == dev][-8:] if graph._CAPACITY is not None else []
            check(lib.mpnn_adam_step(len(ps), ptr_array(ps), ptr_array(grads), ptr_array(ms), ptr_array(vs), numel,
                                     ptr(gs["step"]), ptr(gs["ticket"]), float(group["lr"]), float(b1), float(b2),
                                     float(group["eps"]), float(group["weight_decay"]), ptr_array(guards), len(guards),
                                     stream()), "adam_step")
        return loss
