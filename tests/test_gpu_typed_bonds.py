"""graph.TypedBonds: bond encoder + adjacency-masked batch norm + edge networks evaluated on the
DISTINCT (bond row, adjacency value) pairs of a batch instead of the dense [B,N,N,ef] tensor.  Every test compares the
row-space result with the SAME modules run on the dense tensor (the path that is pinned to the reference's golden
vectors in test_gpu_parity.py)."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

TOL_OUT = 1e-4
TOL_GRAD = 1e-3


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _affinity_batch(B, dev, weighted=False, seed=0):
    from mpnn_b200 import synthetic
    batch = synthetic.make_batch("affinity", B=B, seed_offset=seed)
    t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    if weighted:   # bond-order-like adjacency values: types must be keyed on (row, adjacency value)
        g = torch.Generator().manual_seed(3)
        w = torch.randint(1, 4, t["adj"].shape, generator=g).float().to(dev)
        w = torch.maximum(w, w.transpose(1, 2))
        t["adj"] = t["adj"] * w
    return t


def test_dense_round_trip_bit_exact(dev):
    from mpnn_b200 import graph
    for weighted in (False, True):
        t = _affinity_batch(16, dev, weighted)
        tb = graph.typed_bonds(t["bfm"], t["adj"])
        assert isinstance(tb, graph.TypedBonds)
        assert tuple(tb.shape) == tuple(t["bfm"].shape)
        assert torch.equal(tb.dense(), t["bfm"])
        # counts: every pair of the dense tensor is accounted for exactly once
        B, N = t["adj"].shape[:2]
        assert float(tb._cnt.sum()) == float(B * N * N)
        # adjacency value per type reproduces the dense adjacency
        el = tb.edge_list(t["adj"])
        ti = el.typed()
        pair = el.edge_dst.long() * N + el.edge_src.long() % N
        adj_t = torch.zeros(B * N * N, device=dev)
        adj_t[pair] = tb._a[ti.uid.long()]
        assert torch.equal(adj_t.view(B, N, N), t["adj"])


def test_stock_modules_stay_in_row_space(dev):
    from mpnn_b200 import graph
    t = _affinity_batch(8, dev)
    tb = graph.typed_bonds(t["bfm"], t["adj"])
    torch.manual_seed(1)
    enc = nn.Sequential(nn.Linear(8, 4, bias=False), nn.Tanh(), nn.Linear(4, 3), nn.ReLU(), nn.Softmax(dim=-1),
                        nn.Dropout(0.0), nn.ELU(), nn.Sigmoid()).to(dev)
    y = enc(tb)
    assert isinstance(y, graph.TypedBonds)
    assert tuple(y.shape) == tuple(t["bfm"].shape[:3]) + (3,)
    ref = enc(t["bfm"])
    assert rel_err(y.dense(), ref) <= 1e-6
    # anything that is not row-wise falls back to the dense tensor and still gives the right answer
    assert torch.allclose(tb * 2.0, t["bfm"] * 2.0)
    assert torch.allclose(torch.sum(tb, dim=(1, 2)), t["bfm"].sum(dim=(1, 2)))
    assert torch.allclose(tb.permute(0, 2, 1, 3), t["bfm"].permute(0, 2, 1, 3))


@pytest.mark.parametrize("kind", ["bn1d_train", "bn1d_eval", "bn"])
@pytest.mark.parametrize("weighted", [False, True])
def test_masked_bn_in_row_space(dev, kind, weighted):
    from mpnn_b200 import graph
    from mpnn_b200.modules import MaskBatchNorm, MaskBatchNorm1d
    t = _affinity_batch(16, dev, weighted)
    torch.manual_seed(2)
    enc = nn.Sequential(nn.Linear(8, 4, bias=False), nn.Tanh(), nn.Linear(4, 2)).to(dev)

    def build():
        if kind == "bn":
            return MaskBatchNorm().to(dev)
        bn = MaskBatchNorm1d(2).to(dev)
        with torch.no_grad():
            bn.weight.copy_(torch.tensor([1.3, 0.7]))
            bn.bias.copy_(torch.tensor([0.2, -0.4]))
            bn.running_mean.copy_(torch.tensor([0.1, -0.2]))
            bn.running_var.copy_(torch.tensor([0.9, 1.4]))
        return bn.train() if kind == "bn1d_train" else bn.eval()

    outs = []
    for typed in (True, False):
        bn = build()
        enc.zero_grad()
        x = graph.typed_bonds(t["bfm"], t["adj"]) if typed else t["bfm"]
        y = bn(enc(x), t["adj"])
        if typed:
            assert isinstance(y, graph.TypedBonds)
            y = y.dense()
        g = torch.Generator().manual_seed(4)
        cot = torch.randn(y.shape, generator=g).to(dev)
        (y * cot).sum().backward()
        grads = [p.grad.clone() for p in enc.parameters()] + [p.grad.clone() for p in bn.parameters()]
        bufs = [b.clone() for b in bn.buffers()]
        outs.append((y.detach(), grads, bufs))
    (y1, g1, b1), (y0, g0, b0) = outs
    assert rel_err(y1, y0) <= TOL_OUT
    # the bias in front of a batch norm has a theoretically zero gradient (both paths return round-off): absolute
    # floor relative to the largest gradient of the stack
    gscale = max(float(b.abs().max()) for b in g0)
    for a, b in zip(g1, g0):
        diff = float((a.double() - b.double()).abs().max())
        assert diff <= TOL_GRAD * float(b.abs().max()) + 2e-5 * gscale, (a, b, gscale)
    for a, b in zip(b1, b0):
        assert rel_err(a.float(), b.float()) <= TOL_OUT


def _encoded_model(dev, steps=3, d=8, typed=True):
    """the UNCHANGED reference model file (tests/ref_models/models/normed_encoded_basic_model.py) with the drop-in bond
    encoder (`typed`: mpnn_b200's RowwiseSequential, as `BondAutoEncoder().encoder`) or a stock nn.Sequential"""
    from mpnn_b200.dropin import reference_model as MessagePassingModel, kaiming_init
    from mpnn_b200.modules import RowwiseSequential
    # seed 1: a well-conditioned draw.  With kaiming-initialised 50-layer trunks of width 16 some draws (317 is one)
    # amplify so much that O(100) gradient entries are round-off in BOTH fp32 paths (tools/diag_typed_bonds.py
    # prints row-space / dense / fp64-oracle errors for several seeds: the row-space path is the closer one on average)
    torch.manual_seed(1)
    ae = nn.Sequential(nn.Linear(30, 15, bias=False), nn.Tanh(), nn.Linear(15, d))
    be = (RowwiseSequential if typed else nn.Sequential)(nn.Linear(8, 4, bias=False), nn.Tanh(), nn.Linear(4, 2))
    mod = MessagePassingModel("normed_encoded", d, 2, d, 1, 16, message_steps=steps, atom_encoder=ae, bond_encoder=be)
    mod.apply(kaiming_init)
    return mod.to(dev).train()


@pytest.mark.parametrize("weighted", [False, True])
def test_encoded_model_row_space_equals_dense(dev, weighted, monkeypatch):
    """config 4's caller loop, TypedBonds on vs off.  Outputs and input gradients must agree tightly.  The parameter
    gradients of this stack are sums with heavy cancellation (every Linear sits in front of a batch norm), and the two
    paths add in a different order (per distinct row vs per edge), so both are judged against the fp64 oracle: the
    row-space path must be within tolerance of it, or at least as close to it as the dense path is."""
    from mpnn_b200 import graph
    from oracle import mpnn_oracle as O
    from golden_util import leaf_sd
    t = _affinity_batch(32, dev, weighted)
    res = []
    for typed in (True, False):
        graph.clear_cache()
        mod = _encoded_model(dev, typed=typed)
        sd0 = {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
        afm = t["afm"].clone().requires_grad_(True)
        out = mod(afm, t["bfm"], t["adj"], t["mask"])
        cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).to(dev)
        (out * cot).sum().backward()
        res.append((out.detach(), afm.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters()
                                                      if p.grad is not None},
                    {k: b.clone() for k, b in mod.named_buffers()}))
    (o1, a1, g1, b1), (o0, a0, g0, b0) = res
    assert rel_err(o1, o0) <= TOL_OUT
    assert rel_err(a1, a0) <= TOL_GRAD
    assert set(g1) == set(g0)
    for k in b0:   # running statistics (means of normalised messages are round-off around zero: absolute floor)
        assert float((b1[k].double() - b0[k].double()).abs().max()) <= TOL_OUT * float(b0[k].abs().max()) + 1e-6, k
    # fp64 oracle
    sd = leaf_sd(sd0, dtype=torch.float64)
    c = {k: v.detach().cpu().double() for k, v in t.items()}
    ref = O.normed_encoded_model(c["afm"], c["bfm"], c["adj"], c["mask"], sd, steps=3, buffers={})
    assert rel_err(o1.cpu(), ref.detach()) <= TOL_OUT
    (ref * cot.cpu().double()).sum().backward()
    gscale = max(float(v.grad.abs().max()) for v in sd.values() if getattr(v, "grad", None) is not None)
    for k in g0:
        if getattr(sd[k], "grad", None) is None:
            continue
        r = sd[k].grad
        e1 = float((g1[k].cpu().double() - r).abs().max())
        e0 = float((g0[k].cpu().double() - r).abs().max())
        assert e1 <= max(TOL_GRAD * float(r.abs().max()) + 1e-6 * gscale, 2.0 * e0), (k, e1, e0, float(r.abs().max()))


def test_reference_model_file_usage(dev):
    """the UNCHANGED reference model file (normed_encoded_basic_model.py:67-72): (a) with the drop-in encoder the bond
    tensor is resolved to row space inside `bebn` without the caller doing anything; (b) a stock encoder handed a
    TypedBonds in place of bfm; both equal the dense evaluation"""
    from mpnn_b200 import graph
    from mpnn_b200 import modules as M
    t = _affinity_batch(16, dev)
    mod = _encoded_model(dev, typed=False).eval()
    auto = _encoded_model(dev, typed=True).eval()
    auto.load_state_dict(mod.state_dict())
    seen = []
    orig = M.MaskBatchNorm1d._typed_forward

    def spy(self, tb, mask):
        seen.append(type(tb).__name__)
        return orig(self, tb, mask)

    with torch.no_grad():
        y_dense = mod(t["afm"], t["bfm"], t["adj"], t["mask"])
        y_typed = mod(t["afm"], graph.typed_bonds(t["bfm"], t["adj"]), t["adj"], t["mask"])
        M.MaskBatchNorm1d._typed_forward = spy
        try:
            y_auto = auto(t["afm"], t["bfm"], t["adj"], t["mask"])
        finally:
            M.MaskBatchNorm1d._typed_forward = orig
    assert rel_err(y_typed, y_dense) <= TOL_OUT
    assert rel_err(y_auto, y_dense) <= TOL_OUT
    assert seen == ["TypedBonds"], "the drop-in encoder did not hand the bond rows to bebn in row space"


def test_wrong_adjacency_is_rejected(dev):
    from mpnn_b200 import graph
    from mpnn_b200.modules import EdgeNetwork, AdjMsgAgg
    t = _affinity_batch(4, dev)
    tb = graph.typed_bonds(t["bfm"], t["adj"])
    net = EdgeNetwork(8, 8, 8).to(dev)
    other = t["adj"].clone()
    with pytest.raises(RuntimeError):
        AdjMsgAgg(1)(net(t["afm"][..., :8].contiguous(), tb), other) + 0.0    # (lazy: evaluated by the first consumer)


def test_encoded_step_is_graph_capturable(dev):
    """with the bond features in row space nothing in config 4 needs the host-side edge count: the whole train step
    replays as one CUDA graph and matches the eager step"""
    from mpnn_b200 import graph, graphs
    t = _affinity_batch(64, dev)
    labels = torch.randn(64, 1, generator=torch.Generator().manual_seed(1)).to(dev)
    losses = {}
    for mode in ("eager", "graph"):
        graph.clear_cache()
        mod = _encoded_model(dev)
        head = nn.Linear(16, 1).to(dev)
        torch.manual_seed(5)
        nn.init.normal_(head.weight, std=0.1)
        params = list(mod.parameters()) + list(head.parameters())
        opt = torch.optim.Adam(params, lr=1e-3, capturable=(mode == "graph"))

        batch = dict(t, labels=labels)

        def step_fn(b):
            graph.clear_cache()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(head(mod(b["afm"], b["bfm"], b["adj"], b["mask"])), b["labels"])
            loss.backward()
            opt.step()
            return loss

        if mode == "eager":
            ls = [float(step_fn(batch)) for _ in range(7)][4:]    # GraphedStep: 3 warm-up steps + 1 dry run before capture
        else:
            gs = graphs.GraphedStep(step_fn, batch, warmup=3)
            ls = [float(gs(batch)) for _ in range(3)]
            gs.check()
        losses[mode] = ls
    assert np.allclose(losses["eager"], losses["graph"], rtol=2e-3, atol=1e-6), losses


@pytest.mark.parametrize("bn1d,training", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("R,F", [(7, 2), (33, 8), (300, 5)])
def test_row_space_batch_norm_kernel(dev, bn1d, training, R, F):
    """functional.TypedRowBNFn (csrc/bn.cu k_row_bn_*: the masked batch norms of mask_batch_norm.py on R distinct rows with
    multiplicities) against the same formulas evaluated by torch autograd in fp64"""
    from mpnn_b200.functional import TypedRowBNFn
    g = torch.Generator().manual_seed(R * 7 + F)
    x = torch.randn(R, F, generator=g) * 2 + 0.5
    a = (torch.rand(R, generator=g) > 0.2).float() * (1 + torch.rand(R, generator=g))
    a[0] = 1.0
    cnt = torch.randint(1, 50, (R,), generator=g).float()
    gamma, beta = torch.rand(F, generator=g) + 0.5, torch.randn(F, generator=g) * 0.1
    rm, rv = torch.randn(F, generator=g) * 0.1, torch.rand(F, generator=g) + 0.5
    cot = torch.randn(R, F, generator=g)
    eps, mom = (1e-5, 0.1) if bn1d else (1e-6, 0.0)

    xd, gd, bd = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ad, cd = a.double().unsqueeze(1), cnt.double().unsqueeze(1)
    Mtot = (cd * ad).sum()
    if bn1d:
        if training:
            mean = (cd * ad * xd).sum(0) / Mtot
            var = (cd * ((xd - mean) * ad) ** 2).sum(0) / Mtot
            yr = (xd - mean) / (var.sqrt() + eps)
        else:
            mean, var = rm.double(), rv.double()
            yr = (xd - mean) / (var ** .5 + eps)
        yr = (gd * yr + bd) * ad
    else:
        mean = (cd * xd).sum(0) / Mtot
        var = (cd * ((xd - mean) * ad) ** 2).sum(0) / Mtot
        yr = (xd - mean) * ad / torch.sqrt(var + eps)
    (yr * cot.double()).sum().backward()

    xg = x.to(dev).requires_grad_(True)
    gg, bg = gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
    rm_d, rv_d = rm.clone().to(dev), rv.clone().to(dev)
    y = TypedRowBNFn.apply(xg, a.to(dev), cnt.to(dev), gg if bn1d else None, bg if bn1d else None,
                           rm_d if bn1d else None, rv_d if bn1d else None, bn1d, training, mom, eps)
    (y * cot.to(dev)).sum().backward()
    assert rel_err(y.cpu(), yr.detach()) <= 1e-5
    assert rel_err(xg.grad.cpu(), xd.grad) <= 2e-4
    if bn1d:
        assert rel_err(gg.grad.cpu(), gd.grad) <= 1e-4 and rel_err(bg.grad.cpu(), bd.grad) <= 1e-4
        if training:
            assert rel_err(rm_d.cpu(), (0.9 * rm.double() + 0.1 * mean.detach())) <= 1e-5
            assert rel_err(rv_d.cpu(), (0.9 * rv.double() + 0.1 * var.detach())) <= 1e-5
