"""Shared-parameter gradient hub of the GRU cell (functional.SharedGradSession): the reference applies ONE GRUCell at
every message-passing step (basic_model.py:50-58); the steps' weight-gradient partials are reduced once.  Compared with
the per-call path (MPNN_B200_SHARED_GRAD_HUB=0 semantics), which is the one pinned to the golden vectors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _run(dev, hub, d, rows, steps, monkeypatch, passes=1, retain=False):
    from mpnn_b200 import modules
    monkeypatch.setattr(modules, "SHARED_GRAD_HUB", hub)
    torch.manual_seed(3)
    upd = modules.GRUUpdate(d, d).to(dev)
    g = torch.Generator().manual_seed(11)
    B, N = 4, rows // 4
    mask = (torch.rand(B, N, 1, generator=g) > 0.2).float().to(dev)
    h0 = (torch.randn(B, N, d, generator=g).to(dev) * mask).requires_grad_(True)
    msgs = [torch.randn(B, N, d, generator=g).to(dev).requires_grad_(True) for _ in range(steps)]
    total = 0.0
    for _ in range(passes):
        h = h0
        for m in msgs:
            h = upd(m, h, mask)
        total = total + (h * h).sum()
    if retain:
        total.backward(retain_graph=True)
        first = {k: p.grad.clone() for k, p in upd.named_parameters()}
        upd.zero_grad()
        h0.grad = None
        for m in msgs:
            m.grad = None
        total.backward()
        for k, p in upd.named_parameters():
            assert rel_err(p.grad, first[k]) <= 1e-6, k
    else:
        total.backward()
    return ({k: p.grad.clone() for k, p in upd.named_parameters()}, h0.grad.clone(), [m.grad.clone() for m in msgs])


@pytest.mark.parametrize("d", [8, 16, 19, 32, 64])
@pytest.mark.parametrize("steps", [1, 3, 6])
def test_hub_matches_per_call_gradients(dev, d, steps, monkeypatch):
    a = _run(dev, True, d, 240, steps, monkeypatch)
    b = _run(dev, False, d, 240, steps, monkeypatch)
    tol = 1e-5 if d <= 32 else 5e-3    # d = 64 runs the tensor-core path (no slab: gradients flow through the hub)
    for k in b[0]:
        assert rel_err(a[0][k], b[0][k]) <= tol, k
    assert rel_err(a[1], b[1]) <= tol
    for x, y in zip(a[2], b[2]):
        assert rel_err(x, y) <= tol


def test_hub_two_forward_passes_and_retained_graph(dev, monkeypatch):
    a = _run(dev, True, 16, 240, 3, monkeypatch, passes=2)
    b = _run(dev, False, 16, 240, 3, monkeypatch, passes=2)
    for k in b[0]:
        assert rel_err(a[0][k], b[0][k]) <= 1e-5, k
    _run(dev, True, 16, 240, 3, monkeypatch, retain=True)


def test_hub_new_session_after_backward_and_after_update(dev, monkeypatch):
    from mpnn_b200 import modules
    monkeypatch.setattr(modules, "SHARED_GRAD_HUB", True)
    torch.manual_seed(0)
    upd = modules.GRUUpdate(16, 16).to(dev)
    opt = torch.optim.SGD(upd.parameters(), lr=0.1)
    mask = torch.ones(2, 8, 1, device=dev)
    x = torch.randn(2, 8, 16, device=dev)
    grads = []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        h = x
        for _ in range(3):
            h = upd(x, h, mask)
        h.sum().backward()
        grads.append(upd.gru_cell.weight_ih.grad.clone())
        if it == 0:
            s0 = upd.gru_cell._session
        opt.step()
    assert upd.gru_cell._session is not s0 and s0.done
    assert not torch.equal(grads[0], grads[1])          # parameters moved: fresh gradients, not stale ones
    with torch.no_grad():                               # no tape: no session, plain call
        upd(x, x, mask)
    for p in upd.parameters():
        p.requires_grad_(False)
    h = upd(x.clone().requires_grad_(True), x, mask)    # frozen parameters: per-call path
    h.sum().backward()


def test_inputs_only_backward_leaves_no_stale_partials(dev):
    """a backward pass that differentiates only the inputs (torch.autograd.grad(loss, [afm])) must not leave
    weight-gradient partials in the shared slab that a later backward would add to its parameter gradients"""
    from mpnn_b200 import modules as M
    torch.manual_seed(3)
    uf = M.GRUUpdate(8, 8).to(dev)
    B, N = 4, 6
    m = [torch.randn(B, N, 8, device=dev) for _ in range(3)]
    mask = (torch.rand(B, N, 1, device=dev) > 0.2).float()

    def forward(h0):
        h = h0
        for t in range(3):
            h = uf(m[t], h, mask)
        return h

    h0 = torch.randn(B, N, 8, device=dev, requires_grad=True)
    want_h = forward(h0)
    uf.zero_grad()
    want_h.pow(2).sum().backward()
    want = {k: p.grad.clone() for k, p in uf.named_parameters()}
    # inputs-only gradient first, then a full backward of a NEW forward pass with the same parameters
    h1 = torch.randn(B, N, 8, device=dev, requires_grad=True)
    torch.autograd.grad(forward(h1).sum(), [h1])
    uf.zero_grad()
    h2 = h0.detach().clone().requires_grad_(True)
    forward(h2).pow(2).sum().backward()
    for k, p in uf.named_parameters():
        assert torch.allclose(p.grad, want[k], rtol=1e-5, atol=1e-6), k


def test_two_forward_passes_one_backward_edge_network(dev):
    """two applications of the same EdgeNetwork before ONE backward: the gradients (summed by autograd) must
    equal the run with the side streams switched off"""
    from mpnn_b200 import functional, graph, modules as M, synthetic
    b = synthetic.make_batch("qm9", B=6)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    torch.manual_seed(2)
    net = M.EdgeNetwork(16, 7, 16).to(dev)
    ma = M.AdjMsgAgg(1)
    res = []
    for side in (True, False):
        functional.SIDE_STREAM_ENABLED = side
        functional.SIDE_STREAM_EAGER = side        # (the lanes are capture-only by default)
        try:
            graph.clear_cache()
            net.zero_grad()
            a = ma(net(t["afm"], t["bfm"]), t["adj"]).materialize()
            graph.clear_cache()
            b2 = ma(net(t["afm"] * 0.5, t["bfm"]), t["adj"]).materialize()
            (a.pow(2).sum() + b2.sum()).backward()
            torch.cuda.synchronize()
            res.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
        finally:
            functional.SIDE_STREAM_ENABLED = True
            functional.SIDE_STREAM_EAGER = False
    for k in res[0]:
        assert torch.equal(res[0][k], res[1][k]), k
