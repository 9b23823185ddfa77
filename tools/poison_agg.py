"""Debug aid: the fused aggregation + GRU path on memory that was filled with NaN before (finds reads of never-written
buffers that a fresh process hides behind zero pages)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpnn_b200 import graph, modules as M, synthetic
from mpnn_b200.dropin import reference_model, kaiming_init

dev = torch.device("cuda:0")
d, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 12)
junk = [torch.full((1 << 22,), float("nan"), device=dev) for _ in range(16)]
junk += [torch.full((n,), float("nan"), device=dev) for n in (100, 1000, 5000, 20000, 100000, 400000) for _ in range(8)]
del junk
b = synthetic.make_batch("zinc", B=B, d=d)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
torch.manual_seed(5)
mod = reference_model("normed", d, 8, d, 1, 2 * d, message_steps=2)
mod.apply(kaiming_init)
with torch.no_grad():
    for net in mod.mfs:
        net.edge_map[net._last_idx].weight.mul_(0.05)
mod = mod.to(dev).train()
for fused in (True, False):
    M.AGG_IN_GRU = fused
    graph.clear_cache()
    mod.zero_grad()
    a = t["afm"].clone().requires_grad_(True)
    out = mod(a, t["bfm"], t["adj"], t["mask"])
    out.sum().backward()
    torch.cuda.synchronize()
    print("fused", fused, "out finite", bool(torch.isfinite(out).all()), "grad finite", bool(torch.isfinite(a.grad).all()),
          float(out.abs().max()))
