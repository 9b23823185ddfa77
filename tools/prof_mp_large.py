"""Kernel breakdown of one message-passing step (fwd + bwd) at BASELINE config 5 sizes, on the real rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mpnn_b200 import functional as Fn, graph, modules as M, synthetic

dev = torch.device("cuda:0")
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
h = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w = dict(bench.WORKLOADS["autoenc"], B=Bn, d=h, out=2 * h, targets=2 * h)
batch = synthetic.make_batch("autoenc", B=Bn, d=h)
t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
body, _ = bench.build_model(w, dev)
el = graph.edge_list_for(t["bfm"], t["adj"])
with torch.no_grad():
    tab, tabT = body.mf._compute_table(el)
tab = tab.detach().requires_grad_(True)
cell = body.uf.gru_cell
ws = [p.detach().clone().requires_grad_(True) for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)]
elc, real, hc, mc = M.compact_nodes(t["afm"], t["mask"], el)
Fn.SIDE_STREAM_ENABLED = False


def step():
    Mm = Fn.TypedMessageTCFn.apply(hc, tab, tabT, elc, True, h, h, os.environ.get("MPNN_B200_AGG_IN_GRU", "1") != "0")
    out = Fn.GRUFn.apply(Mm, hc, mc, ws[0], ws[1], ws[2], ws[3], None)
    torch.autograd.grad(out, ws + [tab], torch.ones_like(out))


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = 0.0
print("rows real %d of %d, edges %d" % (hc.shape[0], t["afm"].shape[0] * t["afm"].shape[1], batch["n_edges"]))
for e in evs:
    d = e.time_range.end - e.time_range.start
    tot += d
    print("%8.1f us  %s" % (d, e.name.replace("(anonymous namespace)::", "")[:100]))
print("sum %.1f us" % tot)
