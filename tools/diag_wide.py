"""Diagnostic: a wide (tensor-core) golden model case, error of every output / gradient against the reference golden,
then module by module against the fp64 oracle on the same inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import Case, leaf_sd, rel_err
from mpnn_b200 import modules as M, graph
from mpnn_b200.dropin import reference_model
from oracle import mpnn_oracle as O

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "model_autoencoder_encode_d64"
case = Case(name)
m = case.meta
d = m["d"]
mod = reference_model("autoencoder" if "autoenc" in name else "basic", d, m["ef"], d, 1, m["out"], message_steps=int(m.get("message_steps", 3)))
mod.load_state_dict(case.sd, strict=True)
mod = mod.to(dev).train()
ins = {k: v.clone().to(dev) for k, v in case.inputs.items()}
afm = ins["afm"].requires_grad_(True)
out = mod(afm, ins["bfm"], ins["adj"], ins["mask"])
print("out", rel_err(out.detach().cpu(), case.out["y"]))
(out * case.cot.to(dev)).sum().backward()
print("d afm", rel_err(afm.grad.cpu(), case.gin["afm"]))
params = dict(mod.named_parameters())
for k, g in case.gsd.items():
    got = params[k].grad.cpu() if params[k].grad is not None else torch.zeros_like(g)
    print("  %-28s %.3e  (|g| %.3e)" % (k, rel_err(got, g), float(g.abs().max())))

# ---- module by module, fp64 oracle on CPU with the same weights --------------------------------------------------
sd = leaf_sd(case.sd, dtype=torch.float64)
c = {k: v.double() for k, v in case.inputs.items()}
print("rows", c["afm"].shape, "edges", int((c["adj"] != 0).sum()))
# message + aggregation
a64 = c["afm"].clone().requires_grad_(True)
agg_ref = O.adj_msg_agg(O.edge_network_pairs(a64, c["bfm"], sd, "mf.", d), c["adj"])
g = torch.Generator().manual_seed(3)
cot = torch.randn(agg_ref.shape, generator=g, dtype=torch.float64)
(agg_ref * cot).sum().backward()
graph.clear_cache()
a32 = ins["afm"].detach().clone().requires_grad_(True)
agg = M.AdjMsgAgg(1)(mod.mf(a32, ins["bfm"]), ins["adj"]).materialize()
(agg * cot.float().to(dev)).sum().backward()
print("message+agg: out %.3e  d afm %.3e" % (rel_err(agg.detach().cpu(), agg_ref.detach()), rel_err(a32.grad.cpu(), a64.grad)))
# GRU
msg64 = agg_ref.detach().clone().requires_grad_(True)
h64 = c["afm"].clone().requires_grad_(True)
gru_ref = O.gru_update(msg64, h64, c["mask"], sd, "uf.")
cot = torch.randn(gru_ref.shape, generator=g, dtype=torch.float64)
(gru_ref * cot).sum().backward()
msg32 = agg_ref.detach().float().to(dev).requires_grad_(True)
h32 = ins["afm"].detach().clone().requires_grad_(True)
mod.zero_grad()
gru = mod.uf(msg32, h32, ins["mask"])
(gru * cot.float().to(dev)).sum().backward()
print("gru: out %.3e  d msg %.3e  d h %.3e  dW_ih %.3e dW_hh %.3e" % (
    rel_err(gru.detach().cpu(), gru_ref.detach()), rel_err(msg32.grad.cpu(), msg64.grad), rel_err(h32.grad.cpu(), h64.grad),
    rel_err(mod.uf.gru_cell.weight_ih.grad.cpu(), sd["uf.gru_cell.weight_ih"].grad),
    rel_err(mod.uf.gru_cell.weight_hh.grad.cpu(), sd["uf.gru_cell.weight_hh"].grad)))
# readout
x64 = torch.cat([gru_ref.detach(), c["afm"]], dim=-1).requires_grad_(True)
ro_ref = O.graph_level_output(x64, c["mask"], sd, "of.")
cot = torch.randn(ro_ref.shape, generator=g, dtype=torch.float64)
(ro_ref * cot).sum().backward()
x32 = x64.detach().float().to(dev).requires_grad_(True)
ro = mod.of(x32, mask=ins["mask"])
(ro * cot.float().to(dev)).sum().backward()
print("readout: out %.3e  d x %.3e" % (rel_err(ro.detach().cpu(), ro_ref.detach()), rel_err(x32.grad.cpu(), x64.grad)))
