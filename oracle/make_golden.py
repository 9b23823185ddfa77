"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/, asserts the oracle agrees

Each fixture stores seeded inputs (`in.*`), the reference module's state_dict (`sd.*`), its forward
output(s) (`out.*`), the cotangent used (`cot`) and the gradients autograd produced on the reference
(`gin.*` for inputs, `gsd.*` for parameters).  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _t(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).clone()
    if grad:
        t.requires_grad_(True)
    return t


def _kaiming(m):
    # reference lipo_basic_model.py:88-97 (init_weights) -- nn.Linear branch only
    if type(m) == torch.nn.Linear:
        torch.nn.init.kaiming_uniform_(m.weight, nonlinearity='relu')
        if m.bias is not None:
            torch.nn.init.constant_(m.bias, 0.0)


def _perturb_biases(module, seed):
    """Zero-initialised biases hide indexing bugs: give every bias/buffer-free zero param a seeded value."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            if p.abs().max() == 0:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def save_case(name, inputs, module_sd, outputs, cot, gin, gsd, meta, dedup_aliases=False):
    arrs = {}
    for k, v in inputs.items():
        arrs["in." + k] = v.detach().numpy()
    first = {}
    aliases = {}
    for k, v in module_sd.items():
        if dedup_aliases and ".0.weight" in k and "edge_map" in k:
            # the 50 tied layers are ONE tensor under 50 keys (edge_network.py:20): store it once, list the aliases
            key = (k.split("edge_map")[0], tuple(v.shape), v.detach().numpy().tobytes())
            if key in first:
                aliases[k] = first[key]
                continue
            first[key] = k
        arrs["sd." + k] = v.detach().numpy()
    if aliases:
        meta = dict(meta, sd_aliases=aliases)
    for k, v in outputs.items():
        arrs["out." + k] = v.detach().numpy()
    if cot is not None:
        arrs["cot"] = cot.numpy()
    for k, v in gin.items():
        arrs["gin." + k] = v.numpy()
    for k, v in gsd.items():
        arrs["gsd." + k] = v.numpy()
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrs)
    print("wrote %-28s %6.1f KB" % (name, os.path.getsize(os.path.join(GOLD, name + ".npz")) / 1024))


def run_case(name, module, fwd, inputs, grad_inputs, meta, seed=0, extra_out=None, dedup_aliases=False):
    """fwd(module, **inputs) -> tensor; records grads of `grad_inputs` and all trainable params."""
    ins = {k: _t(v, k in grad_inputs) for k, v in inputs.items()}
    sd0 = {k: v.clone() for k, v in module.state_dict().items()}
    out = fwd(module, **ins)
    g = torch.Generator().manual_seed(1000 + seed)
    cot = torch.randn(out.shape, generator=g)
    (out * cot).sum().backward()
    gin = {k: ins[k].grad.clone() for k in grad_inputs}
    gsd = {}
    for k, p in module.named_parameters():
        if p.requires_grad:
            gsd[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).clone()
    outs = {"y": out}
    if extra_out:
        outs.update(extra_out(module))
    if dedup_aliases:   # gradients of the aliased keys are one tensor too
        seen, keep = {}, {}
        for k, v in gsd.items():
            key = (tuple(v.shape), v.numpy().tobytes()) if (".0.weight" in k and "edge_map" in k) else k
            if key in seen:
                continue
            seen[key] = k
            keep[k] = v
        gsd = keep
    save_case(name, ins, sd0, outs, cot, gin, gsd, meta, dedup_aliases=dedup_aliases)
    return ins, sd0, out, cot, gin, gsd


def make_bilinear(ref):
    """BiLiniearEdgeNetwork: ef = nf**3, parameter-free"""
    sys.path.insert(0, os.path.dirname(HERE))
    from mpnn_b200 import synthetic
    nf = 3
    batch = synthetic.small_batch(B=3, n_lo=2, n_hi=6, afm_width=nf, ef=nf ** 3, seed=13, weighted_adj=True)
    m = ref.BiLiniearEdgeNetwork(nf, nf ** 3, nf)
    run_case("msg_BiLiniearEdgeNetwork", m, lambda mod, afm, bfm, adj: mod(afm, bfm),
             dict(afm=batch["afm"], bfm=batch["bfm"], adj=batch["adj"]), ("afm", "bfm"),
             dict(cls="BiLiniearEdgeNetwork", nf=nf, ef=nf ** 3, mf=nf))


def make_round2(ref):
    """Fixtures added in round 2 (the round-1 files are left untouched): the exported LSTM cell, Set2Vec with a
    caller-supplied initial state, the per-atom readout + BASELINE config 4's ecfp model, and compositions at widths the
    tensor-core kernels serve (d = 64 / 40) plus config 3's shape (d = 32, Set2Vec x 100)."""
    sys.path.insert(0, os.path.dirname(HERE))
    from mpnn_b200 import synthetic
    torch.set_num_threads(4)
    # LSTMCellHidden.forward (set2vec.py:68-75)
    rs = np.random.RandomState(41)
    torch.manual_seed(317)
    m = ref.LSTMCellHidden(12, 6)
    _perturb_biases(m, 15)
    run_case("readout_LSTMCellHidden", m, lambda mod, hprev, cprev: torch.cat(mod(hprev, cprev), dim=1),
             dict(hprev=rs.normal(size=(5, 12)).astype(np.float32), cprev=rs.normal(size=(5, 6)).astype(np.float32)),
             ("hprev", "cprev"), dict(cls="LSTMCellHidden", hd=12, cd=6))
    # Set2Vec with explicit mprev / cprev (set2vec.py:110-116)
    batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=6, ef=3, seed=24)
    torch.manual_seed(317)
    m = ref.Set2Vec(3, 99, time_steps=5)
    _perturb_biases(m, 10)
    run_case("readout_Set2Vec_init", m,
             lambda mod, input_set, mask, mprev, cprev: mod(input_set, mask=mask, mprev=mprev, cprev=cprev),
             dict(input_set=batch["afm"] * batch["mask"], mask=batch["mask"],
                  mprev=rs.normal(size=(4, 6)).astype(np.float32), cprev=rs.normal(size=(4, 6)).astype(np.float32)),
             ("input_set", "mprev", "cprev"), dict(cls="Set2Vec", nf=3, steps=5))
    # per-atom readout (graph_level_output.py:46)
    batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=10, ef=3, seed=23)
    torch.manual_seed(317)
    m = ref.GraphLevelOutputD(5, 7)
    _perturb_biases(m, 9)
    run_case("readout_GraphLevelOutputAtoms", m, lambda mod, input_set, mask: mod(input_set, mask=mask),
             dict(input_set=batch["afm"], mask=batch["mask"]), ("input_set",),
             dict(cls="GraphLevelOutputAtoms", nf=5, out=7))
    # BASELINE config 4: normed_encoded_basic_model_ecfp (no ma_bn, obn on the per-atom readout)
    rawb = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=30, ef=8, seed=32)
    torch.manual_seed(317)
    mod = ref.normed_encoded_basic_model_ecfp.BasicModel(
        8, 2, 8, 1, 5, message_func=ref.EdgeNetworkD, message_opts={}, agg_opts={}, update_opts={},
        readout_func=ref.GraphLevelOutputD, readout_opts={}, message_steps=2,
        atom_encoder=ref.AtomAutoEncoder().encoder, bond_encoder=ref.BondAutoEncoder().encoder)
    mod.apply(_kaiming)
    _perturb_biases(mod, 14)
    mod.train()
    run_case("model_normed_encoded_ecfp", mod, lambda mm, afm, bfm, adj, mask: mm(afm, bfm, adj, mask),
             dict(afm=rawb["afm"], bfm=rawb["bfm"], adj=rawb["adj"], mask=rawb["mask"]), ("afm",),
             dict(cls="normed_encoded_basic_model_ecfp.BasicModel", d=8, ef=2, out=5, steps=2),
             extra_out=lambda mm: {"obn.running_mean": mm.obn.running_mean.clone(),
                                   "obn.running_var": mm.obn.running_var.clone(),
                                   "bebn.running_mean": mm.bebn.running_mean.clone(),
                                   "bebn.running_var": mm.bebn.running_var.clone()})

    # ---- tensor-core widths (d > 32) and config 3's shape --------------------------------------------------------
    def tier_d(name, ctor, d, ef, out, fwd_name="forward", seed=33, B=3, n_hi=7, **kw):
        batch = synthetic.small_batch(B=B, n_lo=2, n_hi=n_hi, afm_width=d, ef=ef, seed=seed)
        torch.manual_seed(317)
        mod = ctor(d, ef, d, 1, out, message_opts={}, agg_opts={}, update_opts={}, **kw)
        mod.apply(_kaiming)
        _perturb_biases(mod, 12)
        mod.train()
        meta = dict(cls=name.split("_d")[0], d=d, ef=ef, out=out)
        meta.update({k: (v if isinstance(v, (int, float, dict)) else str(v)) for k, v in kw.items()})
        run_case(name, mod, lambda mm, afm, bfm, adj, mask: getattr(mm, fwd_name)(afm, bfm, adj, mask),
                 dict(afm=batch["afm"], bfm=batch["bfm"], adj=batch["adj"], mask=batch["mask"]), ("afm",), meta,
                 dedup_aliases=True)

    tier_d("model_autoencoder_encode_d64", ref.basic_graph_autoencoder.Encoder, 64, 8, 128, fwd_name="encode",
           message_func=ref.EdgeNetworkD, message_steps=3, readout_opts={})
    tier_d("model_basic_d40", ref.basic_model.BasicModel, 40, 7, 24, message_func=ref.EdgeNetworkD, message_steps=3,
           readout_opts={})
    tier_d("model_att_d32", ref.att_model.BasicModel, 32, 8, 9, message_func=ref.AttEdgeNetworkD,
           message_agg_func=ref.AdjMsgAgg, message_steps=3, readout_opts={"time_steps": 100})


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "bilinear":     # only the fixture added after the first freeze
        make_bilinear(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "round2":       # fixtures added in round 2 (see make_round2)
        make_round2(ref)
        return
    sys.path.insert(0, os.path.dirname(HERE))
    from mpnn_b200 import synthetic  # host-side numpy only

    torch.set_num_threads(4)

    # ---------------- message functions --------------------------------------------------------
    shapes = [  # (tag, nf, ef, mf, init)
        ("g1", 6, 4, 6, "kaiming"),       # one growth layer, P=16
        ("g1d", 6, 4, 6, "default"),      # default init: trunk decays to ~0, output ~ last bias
        ("rect", 5, 3, 3, "kaiming"),     # nf != mf, P=9
        ("g2", 8, 2, 8, "kaiming"),       # two growth layers 2->4->16
        ("g0", 2, 3, 2, "kaiming"),       # no growth layer, P=ef=3
    ]
    for tag, nf, ef, mf, init in shapes:
        batch = synthetic.small_batch(B=3, n_lo=1, n_hi=6, afm_width=nf, ef=ef, seed=11, weighted_adj=True)
        for cls_name, ocls in (("EdgeNetwork", ref.EdgeNetwork), ("EdgeNetworkD", ref.EdgeNetworkD)):
            torch.manual_seed(317)
            m = ocls(nf, ef, mf)
            if init == "kaiming":
                m.apply(_kaiming)
            _perturb_biases(m, 5)
            run_case("msg_%s_%s" % (cls_name, tag), m, lambda mod, afm, bfm: mod(afm, bfm),
                     dict(afm=batch["afm"], bfm=batch["bfm"]), ("afm", "bfm"),
                     dict(cls=cls_name, nf=nf, ef=ef, mf=mf, init=init))
    for tag, nf, ef, mf in (("g1", 6, 4, 6), ("rect", 5, 3, 3)):
        batch = synthetic.small_batch(B=3, n_lo=1, n_hi=6, afm_width=nf, ef=ef, seed=12, weighted_adj=True)
        torch.manual_seed(317)
        m = ref.AttEdgeNetworkD(nf, ef, mf)
        m.apply(_kaiming)
        _perturb_biases(m, 6)
        run_case("msg_AttEdgeNetworkD_%s" % tag, m, lambda mod, afm, bfm: mod(afm, bfm),
                 dict(afm=batch["afm"], bfm=batch["bfm"]), ("afm", "bfm"),
                 dict(cls="AttEdgeNetworkD", nf=nf, ef=ef, mf=mf, init="kaiming"))

    # GGNN (next row): integer bond types
    rs = np.random.RandomState(3)
    B, N, nf, mf, nt = 2, 5, 4, 4, 3
    bt = rs.randint(0, nt + 1, size=(B, N, N)).astype(np.int64)
    torch.manual_seed(317)
    m = ref.GGNNMsgPass(nf, nt, mf)
    m.init_weights()
    _perturb_biases(m, 7)
    afm = rs.normal(size=(B, N, nf)).astype(np.float32)
    ins = dict(afm=_t(afm, True), bfm=_t(bt))
    out = m(ins["afm"], ins["bfm"])
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * cot).sum().backward()
    save_case("msg_GGNNMsgPass", ins, dict(m.state_dict()), {"y": out}, cot, {"afm": ins["afm"].grad},
              {k: p.grad for k, p in m.named_parameters() if p.grad is not None},
              dict(cls="GGNNMsgPass", nf=nf, ef=nt, mf=mf))

    make_bilinear(ref)

    # ---------------- aggregators (stand-alone, dense [B,N,N,mf]) -------------------------------
    rs = np.random.RandomState(4)
    B, N, mf = 3, 6, 5
    msgs = rs.normal(size=(B, N, N, mf)).astype(np.float32)
    adjw = (rs.rand(B, N, N) < 0.4).astype(np.float32) * rs.uniform(0.5, 2.0, size=(B, N, N)).astype(np.float32)
    for cls_name in ("AdjMsgAgg", "WAdjMsgAgg", "AttMsgAgg"):
        torch.manual_seed(317)
        m = getattr(ref, cls_name)(1)
        run_case("agg_%s" % cls_name, m, lambda mod, messages, adj: mod(messages, adj),
                 dict(messages=msgs, adj=adjw), ("messages", "adj"), dict(cls=cls_name))

    # ---------------- GRU update ---------------------------------------------------------------
    for tag, d in (("d6", 6), ("d22", 22)):
        batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=d, ef=3, seed=21)
        rs = np.random.RandomState(5)
        torch.manual_seed(317)
        m = ref.GRUUpdate(d, d)
        _perturb_biases(m, 8)
        run_case("gru_%s" % tag, m, lambda mod, messages, node_states, mask: mod(messages, node_states, mask),
                 dict(messages=rs.normal(size=batch["afm"].shape).astype(np.float32),
                      node_states=batch["afm"], mask=batch["mask"]),
                 ("messages", "node_states"), dict(cls="GRUUpdate", d=d))

    # ---------------- masked batch norms -------------------------------------------------------
    batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=5, ef=3, seed=22)
    x = batch["afm"] * batch["mask"]
    m = ref.MaskBatchNorm()
    run_case("bn_MaskBatchNorm", m, lambda mod, tensor, mask: mod(tensor, mask),
             dict(tensor=x, mask=batch["mask"]), ("tensor",), dict(cls="MaskBatchNorm"))
    # the 4-D use: bfm with adj as the mask (batch_norm_graph_wrapper.py:14)
    run_case("bn_MaskBatchNorm_bfm", m, lambda mod, tensor, mask: mod(tensor, mask),
             dict(tensor=batch["bfm"] * batch["adj"][..., None], mask=batch["adj"]), ("tensor",),
             dict(cls="MaskBatchNorm"))
    torch.manual_seed(317)
    m = ref.MaskBatchNorm1d(5)
    with torch.no_grad():
        m.weight.copy_(torch.randn(5) * 0.5 + 1)
        m.bias.copy_(torch.randn(5) * 0.1)
    m.train()
    xr = np.random.RandomState(6).normal(size=x.shape).astype(np.float32)  # NOT zero at padding on purpose
    run_case("bn_MaskBatchNorm1d_train", m, lambda mod, tensor, mask: mod(tensor, mask),
             dict(tensor=xr, mask=batch["mask"]), ("tensor",), dict(cls="MaskBatchNorm1d", mode="train"),
             extra_out=lambda mod: {"running_mean": mod.running_mean.clone(), "running_var": mod.running_var.clone()})
    m.eval()
    m.zero_grad()
    run_case("bn_MaskBatchNorm1d_eval", m, lambda mod, tensor, mask: mod(tensor, mask),
             dict(tensor=xr, mask=batch["mask"]), ("tensor",), dict(cls="MaskBatchNorm1d", mode="eval"))

    # ---------------- readouts -----------------------------------------------------------------
    batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=10, ef=3, seed=23)
    torch.manual_seed(317)
    m = ref.GraphLevelOutput(5, 7)
    _perturb_biases(m, 9)
    run_case("readout_GraphLevelOutput", m, lambda mod, input_set, mask: mod(input_set, mask=mask),
             dict(input_set=batch["afm"], mask=batch["mask"]), ("input_set",), dict(cls="GraphLevelOutput", nf=5, out=7))
    m.zero_grad()
    run_case("readout_GraphLevelOutput_nomask", m, lambda mod, input_set: mod(input_set),
             dict(input_set=batch["afm"]), ("input_set",), dict(cls="GraphLevelOutput", nf=5, out=7))
    for tag, steps in (("s7", 7), ("s100", 100)):
        batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=6, ef=3, seed=24)
        torch.manual_seed(317)
        m = ref.Set2Vec(3, 99, time_steps=steps)
        _perturb_biases(m, 10)
        run_case("readout_Set2Vec_%s" % tag, m, lambda mod, input_set, mask: mod(input_set, mask=mask),
                 dict(input_set=batch["afm"] * batch["mask"], mask=batch["mask"]), ("input_set",),
                 dict(cls="Set2Vec", nf=3, steps=steps))

    # ---------------- compositions -------------------------------------------------------------
    # Tier-H: lipo (the only model that runs unmodified at HEAD)
    d, ef = 7, 4
    batch = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=d, ef=ef, seed=31)
    torch.manual_seed(317)
    m = ref.lipo_basic_model.BasicModel(d, ef, d, 1, 9, message_opts={}, agg_opts={}, update_opts={}, readout_opts={},
                                        message_steps=3)
    m.apply(ref.lipo_basic_model.BasicModel.init_weights)
    _perturb_biases(m, 11)
    m.train()
    run_case("model_lipo", m, lambda mod, afm, bfm, adj, mask: mod(afm, bfm, adj, mask),
             dict(afm=batch["afm"], bfm=batch["bfm"], adj=batch["adj"], mask=batch["mask"]), ("afm",),
             dict(cls="lipo_basic_model.BasicModel", d=d, ef=ef, out=9, steps=3),
             extra_out=lambda mod: {"bn.running_mean": mod.bn.running_mean.clone(),
                                    "bn.running_var": mod.bn.running_var.clone(),
                                    "ma_bn.running_mean": mod.ma_bn.running_mean.clone(),
                                    "ma_bn.running_var": mod.ma_bn.running_var.clone()})

    # Tier-D: unchanged reference model files + message_func=EdgeNetworkD / AttEdgeNetworkD
    def tier_d(name, ctor, fwd_name="forward", **kw):
        torch.manual_seed(317)
        mod = ctor(d, ef, d, 1, 9, message_opts={}, agg_opts={}, update_opts={}, readout_opts={}, **kw)
        mod.apply(_kaiming)
        _perturb_biases(mod, 12)
        mod.train()
        run_case(name, mod, lambda mm, afm, bfm, adj, mask: getattr(mm, fwd_name)(afm, bfm, adj, mask),
                 dict(afm=batch["afm"], bfm=batch["bfm"], adj=batch["adj"], mask=batch["mask"]), ("afm",),
                 dict(cls=name, d=d, ef=ef, out=9, **{k: str(v) for k, v in kw.items()}))

    tier_d("model_basic", ref.basic_model.BasicModel, message_func=ref.EdgeNetworkD, message_steps=3)
    tier_d("model_normed_basic", ref.normed_basic_model.BasicModel, message_func=ref.EdgeNetworkD, message_steps=2)
    tier_d("model_autoencoder_encode", ref.basic_graph_autoencoder.Encoder, fwd_name="encode",
           message_func=ref.EdgeNetworkD, message_steps=2)


    for agg_name in ("AdjMsgAgg", "AttMsgAgg"):
        torch.manual_seed(317)
        mod = ref.att_model.BasicModel(d, ef, d, 1, 9, message_func=ref.AttEdgeNetworkD, message_opts={},
                                       message_agg_func=getattr(ref, agg_name), agg_opts={}, update_opts={},
                                       message_steps=2, readout_opts={"time_steps": 6})
        mod.apply(_kaiming)
        _perturb_biases(mod, 13)
        run_case("model_att_%s" % agg_name, mod, lambda mm, afm, bfm, adj, mask: mm(afm, bfm, adj, mask),
                 dict(afm=batch["afm"], bfm=batch["bfm"], adj=batch["adj"], mask=batch["mask"]), ("afm",),
                 dict(cls="att_model.BasicModel", agg=agg_name, d=d, ef=ef, steps=2, s2v_steps=6))

    # encoders + BN1d everywhere (normed_encoded_basic_model), raw 30/8 features -> 8/2
    rawb = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=30, ef=8, seed=32)
    torch.manual_seed(317)
    mod = ref.normed_encoded_basic_model.BasicModel(
        8, 2, 8, 1, 5, message_func=ref.EdgeNetworkD, message_opts={}, agg_opts={}, update_opts={}, readout_opts={},
        message_steps=2, atom_encoder=ref.AtomAutoEncoder().encoder, bond_encoder=ref.BondAutoEncoder().encoder)
    mod.apply(_kaiming)
    _perturb_biases(mod, 14)
    mod.train()
    run_case("model_normed_encoded", mod, lambda mm, afm, bfm, adj, mask: mm(afm, bfm, adj, mask),
             dict(afm=rawb["afm"], bfm=rawb["bfm"], adj=rawb["adj"], mask=rawb["mask"]), ("afm",),
             dict(cls="normed_encoded_basic_model.BasicModel", d=8, ef=2, out=5, steps=2))

    # the layout contract: collate_2d_graphs (data_loader.py:50-70) on ragged graphs
    graphs = synthetic.make_graphs(5, ("uniform", 2, 9), 6, 5, nafm_width=3, seed=99)
    G2 = type("G", (), {})
    objs = []
    for i, g in enumerate(graphs):
        o = G2()
        o.afm, o.nafm, o.bfm, o.adj, o.label = g["afm"], g["nafm"], g["bfm"], g["adj"], float(i)
        objs.append(o)
    coll = ref.collate_2d_graphs(objs)
    mine = synthetic.collate(graphs)
    for k in ("afm", "nafm", "bfm", "adj", "mask"):
        assert np.array_equal(coll[k].numpy(), mine[k]), k
    np.savez_compressed(os.path.join(GOLD, "layout_collate.npz"),
                        **{k: coll[k].numpy() for k in ("afm", "nafm", "bfm", "adj", "mask")},
                        sizes=np.array([g["afm"].shape[0] for g in graphs]))
    print("collate_2d_graphs == synthetic.collate : OK")
    make_round2(ref)


if __name__ == "__main__":
    main()
