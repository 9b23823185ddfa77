"""The fused T-step kernel (csrc/chain.cu, functional.ChainFn) behind the unchanged reference model loops:
fused chain == link-by-link per-module kernels == CPU oracle; bit-reproducible; fallbacks."""
import pytest
import torch
from torch import nn

from golden_util import leaf_sd, rel_err

pytestmark = pytest.mark.gpu

TOL_OUT = 1e-4
TOL_GRAD = 1e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpnn_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _batch(name, B, dev, d=None):
    from mpnn_b200 import synthetic
    b = synthetic.make_batch(name, B=B, d=d) if d else synthetic.make_batch(name, B=B)
    return {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}


def _model(variant, dev, d, ef, out, T, seed=317, **kw):
    from mpnn_b200.dropin import reference_model, kaiming_init
    torch.manual_seed(seed)
    mod = reference_model(variant, d, ef, d, 1, out, message_steps=T, **kw)
    mod.apply(kaiming_init)
    return mod.to(dev).train()


def _run(mod, t, fused, monkeypatch, afm_grad=True):
    from mpnn_b200 import functional, graph
    monkeypatch.setattr(functional, "CHAIN_ENABLED", fused)
    graph.clear_cache()
    calls = []
    orig = functional.ChainFn.forward

    def spy(ctx, *a):
        calls.append(len(a[4]))
        return orig(ctx, *a)

    monkeypatch.setattr(functional.ChainFn, "forward", staticmethod(spy))
    mod.zero_grad()
    afm = t["afm"].clone().requires_grad_(afm_grad)
    out = mod(afm, t["bfm"], t["adj"], t["mask"])
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(out.device)
    (out * cot).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
    bufs = {k: b.clone() for k, b in mod.named_buffers()}
    return out.detach(), (afm.grad.clone() if afm_grad else None), grads, bufs, calls


def _compare(a, b, tol_out=1e-5, tol_grad=1e-4):
    (o1, a1, g1, b1, _), (o0, a0, g0, b0, _) = a, b
    assert rel_err(o1, o0) <= tol_out
    if a0 is not None:
        assert rel_err(a1, a0) <= tol_grad
    assert set(g1) == set(g0)
    gscale = max(float(v.abs().max()) for v in g0.values())
    for k in g0:
        diff = float((g1[k].double() - g0[k].double()).abs().max())
        assert diff <= tol_grad * float(g0[k].abs().max()) + 1e-6 * gscale, k
    for k in b0:
        if b0[k].dtype.is_floating_point:
            assert float((b1[k] - b0[k]).abs().max()) <= 1e-5 * float(b0[k].abs().max()) + 1e-7, k


@pytest.mark.parametrize("variant,T,d,B", [("normed", 3, 16, 32), ("normed", 2, 7, 5), ("basic", 3, 16, 32),
                                            ("autoencoder", 3, 22, 16), ("normed", 8, 32, 12), ("basic", 1, 8, 3)])
def test_fused_chain_equals_link_by_link(dev, variant, T, d, B, monkeypatch):
    """the persistent step kernel against the per-module kernels of round 1 (which are pinned to the reference's
    goldens): outputs, input gradients, every parameter gradient"""
    t = _batch("qm9", B, dev, d=None)
    if d != 16:
        t["afm"] = torch.randn(t["afm"].shape[0], t["afm"].shape[1], d, generator=torch.Generator().manual_seed(1)
                               ).to(dev) * t["mask"]
    mod = _model(variant, dev, d, 7, 24, T)
    fused = _run(mod, t, True, monkeypatch)
    plain = _run(mod, t, False, monkeypatch)
    assert fused[4] == [1 if variant == "autoencoder" else T], "the fused kernel did not serve the loop"
    assert plain[4] == []
    _compare(fused, plain)


def test_fused_chain_matches_oracle_at_config2(dev, monkeypatch):
    """BASELINE config 2 (normed_basic_model, d=16, T=3) at B=32 against the CPU oracle"""
    from mpnn_b200 import synthetic
    from oracle import mpnn_oracle as O
    batch = synthetic.make_batch("qm9", B=32)
    t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    mod = _model("normed", dev, 16, 7, 64, 3)
    sd = leaf_sd({k: v.detach().cpu().clone() for k, v in mod.state_dict().items()})
    out, dafm, grads, _, calls = _run(mod, t, True, monkeypatch)
    assert calls == [3]
    c = {k: torch.from_numpy(batch[k]) for k in ("afm", "bfm", "adj", "mask")}
    a = c["afm"].clone().requires_grad_(True)
    ref = O.normed_basic_model(a, c["bfm"], c["adj"], c["mask"], sd, steps=3)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    (ref * cot).sum().backward()
    assert rel_err(out.cpu(), ref.detach()) <= TOL_OUT
    assert rel_err(dafm.cpu(), a.grad) <= TOL_GRAD
    gscale = max(float(v.grad.abs().max()) for v in sd.values() if getattr(v, "grad", None) is not None)
    for k, g in grads.items():
        if sd[k].grad is None:
            continue
        diff = float((g.cpu().double() - sd[k].grad.double()).abs().max())
        assert diff <= TOL_GRAD * float(sd[k].grad.abs().max()) + 1e-6 * gscale, k


@pytest.mark.parametrize("training", [True, False])
def test_fused_chain_bn1d_ecfp(dev, training, monkeypatch):
    """BASELINE config 4's loop (normed_encoded_basic_model_ecfp.py:67-69: MaskBatchNorm1d after every GRU, encoder
    outputs as differentiable inputs) fused vs link by link, training and eval statistics"""
    from mpnn_b200 import modules as M, synthetic
    b = synthetic.make_batch("affinity", B=24)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    mod = _model("normed_encoded_ecfp", dev, 8, 2, 16, 3, seed=1, readout_func=M.GraphLevelOutputAtoms,
                 atom_encoder=M.AtomAutoEncoder().encoder, bond_encoder=M.BondAutoEncoder().encoder)
    with torch.no_grad():
        for name, p in mod.named_parameters():     # non-trivial affine parameters
            if name.endswith("bn0.weight") or name.endswith("bn1.weight") or name.endswith("bn2.weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            if name.endswith("bn0.bias") or name.endswith("bn1.bias") or name.endswith("bn2.bias"):
                p.copy_(torch.randn_like(p) * 0.1)
    # warm running statistics so eval mode has something to normalise with
    mod.train()
    with torch.no_grad():
        mod(t["afm"], t["bfm"], t["adj"], t["mask"])
    mod.train(training)
    state = {k: v.clone() for k, v in mod.state_dict().items()}
    fused = _run(mod, t, True, monkeypatch, afm_grad=False)
    mod.load_state_dict(state)
    plain = _run(mod, t, False, monkeypatch, afm_grad=False)
    assert fused[4] == [3] and plain[4] == []
    _compare(fused, plain, tol_out=2e-5, tol_grad=2e-3)


def test_fused_chain_is_bit_reproducible(dev, monkeypatch):
    t = _batch("qm9", 64, dev)
    mod = _model("normed", dev, 16, 7, 64, 3)
    a = _run(mod, t, True, monkeypatch)
    b = _run(mod, t, True, monkeypatch)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k


def test_chain_falls_back(dev, monkeypatch):
    """more steps than the kernel holds (T > 8): the first 8 fuse, the rest runs link by link; an intermediate state
    that is consumed mid-loop is evaluated there and the rest of the loop still fuses"""
    from mpnn_b200 import modules as M
    t = _batch("qm9", 6, dev)
    mod = _model("normed", dev, 16, 7, 12, 9)
    r = _run(mod, t, True, monkeypatch)
    assert r[4] == [8] and bool(torch.isfinite(r[0]).all())
    _compare(r, _run(mod, t, False, monkeypatch), tol_out=2e-5, tol_grad=5e-4)
    # a consumer in the middle of the loop evaluates the links so far; the rest still fuses
    mf, ma, uf, bn = M.EdgeNetwork(16, 7, 16).to(dev), M.AdjMsgAgg(1), M.GRUUpdate(16, 16).to(dev), M.MaskBatchNorm()
    h1 = bn(uf(ma(mf(t["afm"], t["bfm"]), t["adj"]), t["afm"], t["mask"]), t["mask"])
    assert isinstance(h1, M.LazyState)
    mid = h1 + 0.0                      # materialises h1
    h2 = bn(uf(ma(mf(t["afm"], t["bfm"], True), t["adj"]), h1, t["mask"]), t["mask"])
    full = torch.cat([h2, t["afm"]], dim=-1)
    assert full.shape[-1] == 32 and torch.equal(mid, h1.materialize())
    # reference semantics of the same two steps, evaluated eagerly
    monkeypatch.setattr(M, "LAZY_CHAIN", False)
    e1 = bn(uf(ma(mf(t["afm"], t["bfm"]), t["adj"]), t["afm"], t["mask"]), t["mask"])
    e2 = bn(uf(ma(mf(t["afm"], t["bfm"], True), t["adj"]), e1, t["mask"]), t["mask"])
    assert rel_err(h2.materialize(), e2) <= 1e-5


def test_fused_chain_large_batch(dev, monkeypatch):
    """many row tiles per CTA (config-4-sized batch): fused == link by link"""
    t = _batch("qm9", 1500, dev)
    mod = _model("normed", dev, 16, 7, 24, 3)
    fused = _run(mod, t, True, monkeypatch)
    plain = _run(mod, t, False, monkeypatch)
    _compare(fused, plain, tol_out=2e-5, tol_grad=5e-4)


def test_captured_step_equals_eager_step(dev):
    """BASELINE config 2's whole train step (fused chain, backward side lanes as graph branches, fused Adam) replayed
    as one CUDA graph against the same step launched eagerly (no side lanes): same losses, same parameters"""
    import numpy as np
    from mpnn_b200 import graph, graphs, optim
    t = _batch("qm9", 48, dev)
    labels = torch.randn(48, 12, generator=torch.Generator().manual_seed(1)).to(dev)
    losses, finals = {}, {}
    for mode in ("eager", "graph"):
        graph.clear_cache()
        mod = _model("normed", dev, 16, 7, 12, 3, seed=11)
        opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)
        batch = dict(t, labels=labels)

        def step_fn(b):
            graph.clear_cache()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(mod(b["afm"], b["bfm"], b["adj"], b["mask"]), b["labels"])
            loss.backward()
            opt.step()
            return loss

        if mode == "eager":
            ls = [float(step_fn(batch)) for _ in range(7)]
        else:
            # a raised overflow flag left behind by an EARLIER capacity session must not gate this one's updates
            graph._CAPTURED_COUNTS.append(torch.tensor([0, 0, 1, 0], dtype=torch.int32, device=dev))
            gs = graphs.GraphedStep(step_fn, batch, warmup=3)
            assert gs.eager_steps == 4
            ls = [float("nan")] * 4 + [float(gs(batch)) for _ in range(3)]
            gs.check()
        losses[mode] = ls[4:]
        finals[mode] = [p.detach().clone() for p in mod.parameters()]
    assert np.allclose(losses["eager"], losses["graph"], rtol=1e-4, atol=1e-7), losses
    for a, b in zip(finals["eager"], finals["graph"]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)


def _padded_batches(dev, n, B, out):
    """n different qm9-shaped batches zero-padded to one common N (a captured step has fixed shapes)"""
    import numpy as np
    from mpnn_b200 import synthetic
    raw = [synthetic.make_batch("qm9", B=B, seed_offset=s) for s in range(n)]
    N = max(b["afm"].shape[1] for b in raw)
    res = []
    for s, b in enumerate(raw):
        n0 = b["afm"].shape[1]
        pad = N - n0
        t = {"afm": np.pad(b["afm"], ((0, 0), (0, pad), (0, 0))),
             "bfm": np.pad(b["bfm"], ((0, 0), (0, pad), (0, pad), (0, 0))),
             "adj": np.pad(b["adj"], ((0, 0), (0, pad), (0, pad))),
             "mask": np.pad(b["mask"], ((0, 0), (0, pad)) + ((0, 0),) * (b["mask"].ndim - 2))}
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in t.items()}
        t["labels"] = torch.randn(B, out, generator=torch.Generator().manual_seed(100 + s)).to(dev)
        res.append(t)
    return res


def test_pipelined_prep_equals_plain_captured_step(dev):
    """GraphedStep(pipeline_prep=True): a batch's compaction / de-duplication / type sort run one replay AHEAD, beside
    the previous batch's train step.  Over a sequence of DIFFERENT batches the losses and the parameters are
    bit-identical to the plain captured step (same kernels on the same edge lists, only scheduled earlier)."""
    from mpnn_b200 import graph, graphs, optim
    seq = _padded_batches(dev, 5, 40, 12)
    order = [0, 1, 2, 3, 4, 2, 0]
    losses, finals = {}, {}
    for mode in ("plain", "pipelined"):
        graph.clear_cache()
        mod = _model("normed", dev, 16, 7, 12, 3, seed=11)
        opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)

        def step_fn(b):
            graph.clear_cache()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(mod(b["afm"], b["bfm"], b["adj"], b["mask"]), b["labels"])
            loss.backward()
            opt.step()
            return loss

        gs = graphs.GraphedStep(step_fn, seq[0], warmup=3, edge_capacity=4096, unique_capacity=64,
                                pipeline_prep=(mode == "pipelined"))
        ls = []
        if mode == "plain":
            for i in order:
                ls.append(float(gs(seq[i])))
        else:
            # the constructor prepared seq[0] (the example batch); every following batch's bfm / adj are announced one
            # replay before its afm / mask / labels are loaded
            for k in range(len(order)):
                gs.load(seq[order[k]])
                gs.load_next(seq[order[min(k + 1, len(order) - 1)]])
                ls.append(float(gs.replay()))
        gs.check()
        losses[mode] = ls
        finals[mode] = [p.detach().clone() for p in mod.parameters()]
    assert losses["plain"] == losses["pipelined"], losses
    for a, b in zip(finals["plain"], finals["pipelined"]):
        assert torch.equal(a, b)
    assert len(set(losses["plain"][:5])) == 5      # the batches really differ


def test_overflow_flag_survives_later_replays(dev):
    """plain captured step: an overflowing batch in the MIDDLE of a replayed sequence skips its update on the device and
    is still reported by the next check(), although later replays rewrite the per-step counts"""
    from mpnn_b200 import graph, graphs, optim
    seq = _padded_batches(dev, 2, 40, 12)
    graph.clear_cache()
    mod = _model("normed", dev, 16, 7, 12, 3, seed=11)
    opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)

    def step_fn(b):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(mod(b["afm"], b["bfm"], b["adj"], b["mask"]), b["labels"])
        loss.backward()
        opt.step()
        return loss

    small = {k: v.clone() for k, v in seq[0].items()}
    small["bfm"][20:] = 0
    small["adj"][20:] = 0
    e_small = int(((small["bfm"] != 0).any(-1) | (small["adj"] != 0)).sum())
    gs = graphs.GraphedStep(step_fn, small, warmup=3, edge_capacity=e_small + 64, unique_capacity=64)
    gs(small)
    gs.check()
    before = [p.detach().clone() for p in mod.parameters()]
    gs(seq[1])                       # overflows: the update is gated off
    for a, b in zip(before, [p.detach().clone() for p in mod.parameters()]):
        assert torch.equal(a, b)
    gs(small)
    gs(small)
    with pytest.raises(RuntimeError, match="capacit"):
        gs.check()
    gs.check()


@pytest.mark.parametrize("pipelined", [False, True])
def test_host_io_graph_equals_plain_captured_step(dev, pipelined):
    """GraphedStep(host_io=True): the H2D of the host's pinned buffers and the device copy into the static inputs are
    nodes of the step's graph.  Over a sequence of different batches fed through the pinned buffers the losses and the
    parameters equal the plain captured step's (with and without the pipelined preprocessing)."""
    from mpnn_b200 import graph, graphs, optim
    seq = _padded_batches(dev, 4, 40, 12)
    order = [0, 1, 2, 3, 1, 0]
    losses, finals = {}, {}
    for mode in ("plain", "host_io"):
        graph.clear_cache()
        mod = _model("normed", dev, 16, 7, 12, 3, seed=11)
        opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)

        def step_fn(b):
            graph.clear_cache()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(mod(b["afm"], b["bfm"], b["adj"], b["mask"]), b["labels"])
            loss.backward()
            opt.step()
            return loss

        if mode == "plain":
            gs = graphs.GraphedStep(step_fn, seq[0], warmup=3, edge_capacity=4096, unique_capacity=64)
            ls = [float(gs(seq[i])) for i in order]
        else:
            gs = graphs.GraphedStep(step_fn, seq[0], warmup=3, edge_capacity=4096, unique_capacity=64,
                                    pipeline_prep=pipelined, host_io=True)
            at = lambda k: seq[order[min(k, len(order) - 1)]]

            def fill(views, batch, keys):
                for k_ in keys:
                    views[k_].copy_(batch[k_])

            if pipelined:   # the staged bfm / adj are one batch ahead of the staged afm / mask / labels
                fill(gs.stage["bonds"][1], at(1), ("bfm", "adj"))
            ls = []
            for k in range(len(order)):
                torch.cuda.synchronize()        # the previous replay has read the pinned buffers
                fill(gs.host["bonds"][1], at(k + 2) if pipelined else at(k + 1), ("bfm", "adj"))
                fill(gs.host["rest"][1], at(k + 1), ("afm", "mask", "labels"))
                ls.append(float(gs.replay()))
        gs.check()
        losses[mode] = ls
        finals[mode] = [p.detach().clone() for p in mod.parameters()]
    assert losses["plain"] == losses["host_io"], losses
    for a, b in zip(finals["plain"], finals["host_io"]):
        assert torch.equal(a, b)


def test_pipelined_prep_overflow_flag_is_sticky(dev):
    """an overflowing batch anywhere in the replayed sequence is reported by check(), and its update is skipped"""
    from mpnn_b200 import graph, graphs, optim
    seq = _padded_batches(dev, 2, 40, 12)
    graph.clear_cache()
    mod = _model("normed", dev, 16, 7, 12, 3, seed=11)
    opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)

    def step_fn(b):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(mod(b["afm"], b["bfm"], b["adj"], b["mask"]), b["labels"])
        loss.backward()
        opt.step()
        return loss

    small = {k: v.clone() for k, v in seq[0].items()}
    small["bfm"][20:] = 0
    small["adj"][20:] = 0        # half of the graphs lose their bonds: far fewer edges than seq[1]
    e_small = int(((small["bfm"] != 0).any(-1) | (small["adj"] != 0)).sum())
    e_big = int(((seq[1]["bfm"] != 0).any(-1) | (seq[1]["adj"] != 0)).sum())
    assert e_big > e_small + 80
    gs = graphs.GraphedStep(step_fn, small, warmup=3, edge_capacity=e_small + 64, unique_capacity=64,
                            pipeline_prep=True)
    gs.load(small)
    gs.load_next(seq[1])
    gs.replay()                      # trains on `small`, prepares seq[1] (overflows)
    before = [p.detach().clone() for p in mod.parameters()]
    gs.load(seq[1])
    gs.load_next(small)
    gs.replay()                      # trains on the truncated seq[1]: the update is gated off
    after = [p.detach().clone() for p in mod.parameters()]
    for a, b in zip(before, after):
        assert torch.equal(a, b)
    gs.load(small)
    gs.replay()                      # a fitting batch again
    with pytest.raises(RuntimeError, match="capacit"):
        gs.check()
    gs.check()                       # cleared by the first check


def test_captured_att_step_equals_eager_step(dev):
    """BASELINE config 3's train step (att_model: AttEdgeNetwork's per-edge gate on edge SLOTS in capacity mode, Set2Vec's
    persistent kernels, fused Adam) replayed as one CUDA graph against the eager step"""
    import numpy as np
    from mpnn_b200 import graph, graphs, modules as M, optim, synthetic
    b = synthetic.make_batch("zinc", B=24)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    d, ef = t["afm"].shape[-1], t["bfm"].shape[-1]
    labels = torch.randn(24, 4 * d, generator=torch.Generator().manual_seed(1)).to(dev)
    losses, finals = {}, {}
    for mode in ("eager", "graph"):
        graph.clear_cache()
        mod = _model("att", dev, d, ef, 4 * d, 3, seed=5, message_func=M.AttEdgeNetwork, readout_func=M.Set2Vec,
                     readout_opts={"time_steps": 10})
        opt = optim.FusedAdam(list(mod.parameters()), lr=1e-3)
        batch = dict(t, labels=labels)

        def step_fn(bb):
            graph.clear_cache()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(mod(bb["afm"], bb["bfm"], bb["adj"], bb["mask"]), bb["labels"])
            loss.backward()
            opt.step()
            return loss

        if mode == "eager":
            ls = [float(step_fn(batch)) for _ in range(7)][4:]
        else:
            gs = graphs.GraphedStep(step_fn, batch, warmup=3)
            ls = [float(gs(batch)) for _ in range(3)]
            gs.check()
        losses[mode] = ls
        finals[mode] = [p.detach().clone() for p in mod.parameters()]
    assert np.allclose(losses["eager"], losses["graph"], rtol=1e-4, atol=1e-7), losses
    for a, c in zip(finals["eager"], finals["graph"]):
        assert torch.allclose(a, c, rtol=2e-3, atol=2e-5)
