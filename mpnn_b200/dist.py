"""Data-parallel plumbing: graphs shard by rank, one gradient all-reduce per step.

One process per GPU (torchrun); `torch.distributed` with NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests.  The path has no forward/inference communication: every message-passing op is per-graph, so
the only exchange is the parameter-gradient sum.  Parameter sets here are small (27 k .. 271 k floats for
the benchmark configs = 0.1 .. 1.1 MB), i.e. latency-bound: ONE flat bucket, one collective per step.
Per-replica semantics for the batch statistics (masked BNs, Set2Vec's batch softmax): each rank's forward
equals the reference run on that rank's local batch.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises the default process group from torchrun's env (RANK/WORLD_SIZE/MASTER_*). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_bounds(n_graphs, rank, world):
    """Contiguous graph range [lo, hi) of this rank (graphs are the independent units of the path)."""
    per = (n_graphs + world - 1) // world
    lo = min(n_graphs, rank * per)
    return lo, min(n_graphs, lo + per)


def shard_batch(batch, rank, world):
    """Slices a padded batch dict by graph and re-pads to the LOCAL maximum size (as collate_2d_graphs would)."""
    n_graphs = batch["afm"].shape[0]
    lo, hi = shard_bounds(n_graphs, rank, world)
    mask = batch["mask"][lo:hi]
    n_local = int(mask.reshape(hi - lo, -1).sum(1).max()) if hi > lo else 0
    n_local = max(n_local, 1)
    out = {}
    for k, v in batch.items():
        if not hasattr(v, "shape") or getattr(v, "ndim", 0) == 0 or v.shape[0] != n_graphs:
            out[k] = v
            continue
        s = v[lo:hi]
        if k in ("afm", "nafm", "mask"):
            s = s[:, :n_local]
        elif k == "bfm":
            s = s[:, :n_local, :n_local]
        elif k == "adj":
            s = s[:, :n_local, :n_local]
        out[k] = s
    return out


class FlatGradAllReduce(object):
    """Sums (then averages) all parameter gradients in one flat bucket, in place."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self._flat = None

    def __call__(self, average=True):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        # the bucket layout is the list of parameters that REQUIRE a gradient -- the same on every rank whatever
        # happened to receive one in this step; a parameter without a gradient contributes zeros (and keeps None if no
        # rank had one: its slice of the sum is then zero and is not written back)
        ps = self.params
        if not ps:
            return
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        if self._flat is None or self._flat.numel() != n or self._flat.device != dev:
            self._flat = torch.empty(n, dtype=torch.float32, device=dev)
        views, o = [], 0
        for p in ps:
            views.append(self._flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        have = [i for i, p in enumerate(ps) if p.grad is not None]
        if len(have) != len(ps):
            self._flat.zero_()
        if have:
            torch._foreach_copy_([views[i] for i in have], [ps[i].grad for i in have])
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        if average:
            self._flat.div_(dist.get_world_size())
        if have:
            torch._foreach_copy_([ps[i].grad for i in have], [views[i] for i in have])
        for i, p in enumerate(ps):
            if p.grad is None and len(have) != len(ps):
                p.grad = views[i].clone()    # another rank may have produced this gradient: take the averaged sum
