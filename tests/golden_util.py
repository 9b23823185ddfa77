"""Helpers shared by the CPU (oracle vs golden) and GPU (CUDA vs golden/oracle) parity tests."""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


class Case(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLD, name + ".npz"))
        self.name = name
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.inputs, self.sd, self.out, self.gin, self.gsd = {}, {}, {}, {}, {}
        for k in z.files:
            head, _, rest = k.partition(".")
            if head == "in":
                self.inputs[rest] = torch.from_numpy(z[k])
            elif head == "sd":
                self.sd[rest] = torch.from_numpy(z[k])
            elif head == "out":
                self.out[rest] = torch.from_numpy(z[k])
            elif head == "gin":
                self.gin[rest] = torch.from_numpy(z[k])
            elif head == "gsd":
                self.gsd[rest] = torch.from_numpy(z[k])
        self.cot = torch.from_numpy(z["cot"]) if "cot" in z.files else None
        for alias, first in self.meta.get("sd_aliases", {}).items():   # the 50 tied layers, stored once
            self.sd[alias] = self.sd[first]


def all_cases(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz"))
                  if not os.path.basename(p).startswith("layout_"))


def leaf_sd(sd, dtype=torch.float32):
    """state_dict -> dict of leaf tensors requiring grad; aliased entries (the 50 tied layers) share ONE leaf."""
    by_bytes, out = {}, {}
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        if ".0.weight" in k and "edge_map" in k:
            key = (k.split("edge_map")[0], v.shape, v.numpy().tobytes())
            if key not in by_bytes:
                by_bytes[key] = v.clone().to(dtype).requires_grad_(True)
            out[k] = by_bytes[key]
        else:
            out[k] = v.clone().to(dtype).requires_grad_(True)
    return out


def oracle_forward(case, ins, sd, buffers=None):
    """Runs the oracle restatement for a golden case. Returns the output tensor."""
    from oracle import mpnn_oracle as O
    m = case.meta
    cls = m["cls"]
    if cls == "EdgeNetwork":
        return O.edge_network_head(ins["afm"], ins["bfm"], sd, "", m["mf"])
    if cls == "EdgeNetworkD":
        return O.edge_network_pairs(ins["afm"], ins["bfm"], sd, "", m["mf"])
    if cls == "AttEdgeNetworkD":
        return O.att_edge_network_pairs(ins["afm"], ins["bfm"], sd, "", m["mf"])
    if cls == "GGNNMsgPass":
        return O.ggnn_msg_pass(ins["afm"], ins["bfm"], sd, "")
    if cls == "BiLiniearEdgeNetwork":
        return O.bilinear_edge_network(ins["afm"], ins["bfm"], m["nf"])
    if cls == "AdjMsgAgg":
        return O.adj_msg_agg(ins["messages"], ins["adj"])
    if cls == "WAdjMsgAgg":
        return O.wadj_msg_agg(ins["messages"], ins["adj"])
    if cls == "AttMsgAgg":
        return O.att_msg_agg(ins["messages"], ins["adj"], sd, "")
    if cls == "GRUUpdate":
        return O.gru_update(ins["messages"], ins["node_states"], ins["mask"], sd, "")
    if cls == "MaskBatchNorm":
        return O.mask_batch_norm(ins["tensor"], ins["mask"])
    if cls == "MaskBatchNorm1d":
        return O.mask_batch_norm_1d(ins["tensor"], ins["mask"], sd, "", training=(m["mode"] == "train"),
                                    buffers=buffers)
    if cls == "GraphLevelOutput":
        return O.graph_level_output(ins["input_set"], ins.get("mask"), sd, "")
    if cls == "GraphLevelOutputAtoms":
        return O.graph_level_output_atoms(ins["input_set"], ins["mask"], sd, "")
    if cls == "LSTMCellHidden":
        return torch.cat(O.lstm_cell_hidden(ins["hprev"], ins["cprev"], sd, ""), dim=1)
    if cls == "Set2Vec":
        return O.set2vec(ins["input_set"], ins["mask"], sd, "", steps=m["steps"], mprev=ins.get("mprev"),
                         cprev=ins.get("cprev"))
    a = (ins["afm"], ins["bfm"], ins["adj"], ins["mask"])
    if cls == "lipo_basic_model.BasicModel":
        return O.lipo_model(*a, sd=sd, steps=m["steps"], buffers=buffers)
    if cls == "model_basic":
        return O.basic_model(*a, sd=sd, steps=int(m.get("message_steps", 3)))
    if cls == "model_normed_basic":
        return O.normed_basic_model(*a, sd=sd, steps=2)
    if cls == "model_autoencoder_encode":
        return O.basic_model(*a, sd=sd, steps=int(m.get("message_steps", 2)), chain_state=False)
    if cls == "model_att":
        return O.att_model(*a, sd=sd, steps=int(m["message_steps"]), s2v_steps=m["readout_opts"]["time_steps"], agg="adj")
    if cls == "normed_encoded_basic_model_ecfp.BasicModel":
        return O.normed_encoded_ecfp_model(*a, sd=sd, steps=m["steps"], buffers=buffers)
    if cls == "att_model.BasicModel":
        return O.att_model(*a, sd=sd, steps=m["steps"], s2v_steps=m["s2v_steps"],
                           agg="adj" if m["agg"] == "AdjMsgAgg" else "att")
    if cls == "normed_encoded_basic_model.BasicModel":
        return O.normed_encoded_model(*a, sd=sd, steps=m["steps"], buffers=buffers)
    raise KeyError(cls)


def run_oracle(case, dtype=torch.float32):
    ins = {}
    for k, v in case.inputs.items():
        t = v.clone()
        if t.dtype.is_floating_point:
            t = t.to(dtype)
            if k in case.gin:
                t.requires_grad_(True)
        ins[k] = t
    sd = leaf_sd(case.sd, dtype)
    buffers = {}
    out = oracle_forward(case, ins, sd, buffers)
    (out * case.cot.to(dtype)).sum().backward()
    gin = {k: ins[k].grad for k in case.gin}
    gsd = {}
    for k in case.gsd:
        g = sd[k].grad
        gsd[k] = g if g is not None else torch.zeros_like(sd[k])
    return out.detach(), gin, gsd, buffers


def rel_err(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
