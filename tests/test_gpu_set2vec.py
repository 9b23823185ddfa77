"""Set2Vec (reference readout/set2vec.py:93-151): the persistent loop kernels (csrc/s2v_persist.cu) against the
per-iteration launches (csrc/set2vec.cu, pinned to the reference's goldens in test_gpu_parity.py) and the CPU oracle."""
import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpnn_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _inputs(B, N, F, dev, seed=3, ragged=True):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(B, N, F, generator=g)
    n = torch.randint(1, N + 1, (B,), generator=g) if ragged else torch.full((B,), N)
    n[0] = N
    mask = (torch.arange(N)[None, :] < n[:, None]).float()[..., None]
    return (X * mask).to(dev), mask.to(dev)


def _run(mod, X, mask, persistent, m0=None, c0=None):
    from mpnn_b200 import functional
    prev = functional.set2vec_persistent(persistent)
    try:
        mod.zero_grad()
        x = X.clone().requires_grad_(True)
        ins = [t.clone().requires_grad_(True) if t is not None else None for t in (m0, c0)]
        out = mod(x, mask, ins[0], ins[1]) if m0 is not None else mod(x, mask)
        cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).to(out.device)
        (out * cot).sum().backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
        extra = [t.grad.clone() if t is not None else None for t in ins]
        return out.detach(), x.grad.clone(), grads, extra
    finally:
        functional.set2vec_persistent(prev)


# (B, N, F/2 = node_features, steps): one graph per CTA, several per CTA (B > #SMs), X not resident (B*N*F too large),
# odd widths, a single graph, the reference's default 100 iterations
SHAPES = [(6, 9, 16, 5), (128, 38, 32, 12), (300, 20, 32, 4), (1100, 30, 32, 3), (7, 13, 11, 6), (1, 5, 4, 3),
          (32, 25, 32, 100), (2000, 12, 8, 3)]


@pytest.mark.parametrize("B,N,nf,steps", SHAPES)
def test_persistent_equals_per_step(dev, B, N, nf, steps):
    from mpnn_b200 import modules as M
    torch.manual_seed(B + steps)
    mod = M.Set2Vec(nf, 99, time_steps=steps).to(dev)
    X, mask = _inputs(B, N, 2 * nf, dev)
    a = _run(mod, X, mask, True)
    b = _run(mod, X, mask, False)
    assert rel_err(a[0], b[0]) <= 2e-5
    assert rel_err(a[1], b[1]) <= 2e-4
    assert set(a[2]) == set(b[2])
    for k in b[2]:
        assert rel_err(a[2][k], b[2][k]) <= 5e-4, k


def test_persistent_initial_state_and_reproducibility(dev):
    """caller-supplied (mprev, cprev) (set2vec.py:111-117) with gradients; two runs are bit-identical"""
    from mpnn_b200 import modules as M
    torch.manual_seed(1)
    B, N, nf, steps = 40, 17, 16, 7
    mod = M.Set2Vec(nf, 99, time_steps=steps).to(dev)
    X, mask = _inputs(B, N, 2 * nf, dev)
    g = torch.Generator().manual_seed(8)
    m0 = torch.randn(B, 2 * nf, generator=g).to(dev)
    c0 = torch.randn(B, 2 * nf, generator=g).to(dev)
    a = _run(mod, X, mask, True, m0, c0)
    b = _run(mod, X, mask, False, m0, c0)
    assert rel_err(a[0], b[0]) <= 2e-5 and rel_err(a[1], b[1]) <= 2e-4
    for k in b[2]:
        assert rel_err(a[2][k], b[2][k]) <= 5e-4, k
    for u, v in zip(a[3], b[3]):
        assert rel_err(u, v) <= 2e-4
    a2 = _run(mod, X, mask, True, m0, c0)
    assert torch.equal(a[0], a2[0]), "forward"
    assert torch.equal(a[1], a2[1]), "dX"
    for k in a[2]:
        assert torch.equal(a[2][k], a2[2][k]), k


def test_persistent_against_oracle(dev):
    """the CPU restatement of set2vec.py on the same weights"""
    from mpnn_b200 import modules as M
    from oracle import mpnn_oracle as O
    torch.manual_seed(2)
    B, N, nf, steps = 24, 15, 12, 20
    mod = M.Set2Vec(nf, 99, time_steps=steps).to(dev)
    X, mask = _inputs(B, N, 2 * nf, dev)
    out, dX, grads, _ = _run(mod, X, mask, True)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    x = X.cpu().clone().requires_grad_(True)
    ref = O.set2vec(x, mask.cpu(), sd, "", steps=steps)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (ref * cot).sum().backward()
    assert rel_err(out.cpu(), ref.detach()) <= 1e-4
    assert rel_err(dX.cpu(), x.grad) <= 1e-3
    for k, g in grads.items():
        assert rel_err(g.cpu(), sd[k].grad) <= 2e-3, k


def test_persistent_without_mask(dev):
    """Set2Vec.forward(input_set) with mask=None (set2vec.py:120: no -1e8 term): persistent == per-iteration"""
    from mpnn_b200 import functional, modules as M
    torch.manual_seed(4)
    mod = M.Set2Vec(8, 99, time_steps=9).to(dev)
    X = torch.randn(20, 11, 16, generator=torch.Generator().manual_seed(1)).to(dev)
    res = []
    for persistent in (True, False):
        prev = functional.set2vec_persistent(persistent)
        try:
            mod.zero_grad()
            x = X.clone().requires_grad_(True)
            out = mod(x)
            out.pow(2).sum().backward()
            torch.cuda.synchronize()
            res.append((out.detach(), x.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters()}))
        finally:
            functional.set2vec_persistent(prev)
    a, b = res
    assert rel_err(a[0], b[0]) <= 2e-5 and rel_err(a[1], b[1]) <= 2e-4
    for k in b[2]:
        assert rel_err(a[2][k], b[2][k]) <= 5e-4, k
