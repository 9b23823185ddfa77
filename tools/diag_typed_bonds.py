"""Diagnostic: parameter gradients of config 4's stack, row-space (TypedBonds) vs dense path vs the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch import nn
from mpnn_b200 import graph, synthetic
from mpnn_b200.modules import RowwiseSequential
from mpnn_b200.dropin import reference_model as MessagePassingModel, kaiming_init
from oracle import mpnn_oracle as O
from golden_util import leaf_sd

dev = torch.device("cuda:0")

def model(seed, tame, typed=True):
    torch.manual_seed(seed)
    ae = nn.Sequential(nn.Linear(30, 15, bias=False), nn.Tanh(), nn.Linear(15, 8))
    be = (RowwiseSequential if typed else nn.Sequential)(nn.Linear(8, 4, bias=False), nn.Tanh(), nn.Linear(4, 2))
    mod = MessagePassingModel("normed_encoded", 8, 2, 8, 1, 16, message_steps=3, atom_encoder=ae, bond_encoder=be)
    mod.apply(kaiming_init)
    if tame:
        with torch.no_grad():
            for mf in mod.mfs:
                mf.edge_map[mf._tied_idx][0].weight.mul_(tame)
    return mod.to(dev).train()

for weighted in (False, True):
    for seed in (317, 1, 2):
        for tame in (0, 0.8):
            batch = synthetic.make_batch("affinity", B=32)
            t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
            if weighted:
                g = torch.Generator().manual_seed(3)
                w = torch.randint(1, 4, t["adj"].shape, generator=g).float().to(dev)
                t["adj"] = t["adj"] * torch.maximum(w, w.transpose(1, 2))
            res = []
            for typed in (True, False):
                graph.clear_cache()
                mod = model(seed, tame, typed)
                sd0 = {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
                out = mod(t["afm"], t["bfm"], t["adj"], t["mask"])
                cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(dev)
                (out * cot).sum().backward()
                res.append({k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None})
            sd = leaf_sd(sd0, dtype=torch.float64)
            c = {k: v.detach().cpu().double() for k, v in t.items()}
            ref = O.normed_encoded_model(c["afm"], c["bfm"], c["adj"], c["mask"], sd, steps=3, buffers={})
            (ref * cot.cpu().double()).sum().backward()
            gscale = max(float(v.grad.abs().max()) for v in sd.values() if getattr(v, "grad", None) is not None)
            print("weighted=%s seed=%d tame=%s gscale=%.3g" % (weighted, seed, tame, gscale))
            worst = []
            for k in res[0]:
                if getattr(sd[k], "grad", None) is None:
                    continue
                r = sd[k].grad
                e1 = float((res[0][k].cpu().double() - r).abs().max())
                e0 = float((res[1][k].cpu().double() - r).abs().max())
                worst.append((max(e1, e0) / (float(r.abs().max()) + 1e-30), k, e1, e0, float(r.abs().max())))
            worst.sort(reverse=True)
            for w_ in worst[:4]:
                print("   %-28s typed_err=%.3g dense_err=%.3g |ref|=%.3g" % w_[1:])
