import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from mpnn_b200 import modules as M, synthetic, graph
from mpnn_b200.dropin import reference_model, kaiming_init
dev = torch.device("cuda:0")
def poison():
    xs = [torch.full((64 * 1024 * 1024,), float("nan"), device=dev) for _ in range(8)]
    # many small blocks too
    ys = [torch.full((n,), float("nan"), device=dev) for n in [1000, 5000, 20000, 100000, 500000] * 40]
    del xs, ys
    torch.cuda.synchronize()
torch.manual_seed(317)
batch = synthetic.make_batch("zinc", B=6)
mod = reference_model("att", 32, 8, 32, 1, 128, message_func=M.AttEdgeNetwork, message_agg_func=M.AdjMsgAgg,
                      message_steps=3, readout_func=M.Set2Vec, readout_opts={"time_steps": 100})
mod.apply(kaiming_init)
mod = mod.to(dev).train()
t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
res = []
for trial in range(3):
    poison()
    graph.clear_cache()
    mod.zero_grad()
    afm = t["afm"].clone().requires_grad_(True)
    out = mod(afm, t["bfm"], t["adj"], t["mask"])
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(dev)
    (out * cot).sum().backward()
    torch.cuda.synchronize()
    bad = [k for k, p in mod.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print("trial", trial, "out finite", bool(torch.isfinite(out).all()), "afm.grad finite", bool(torch.isfinite(afm.grad).all()), "bad params", bad[:6])
    res.append(afm.grad.clone())
print("trial diffs", float((res[0] - res[1]).abs().max()), float((res[1] - res[2]).abs().max()))
