"""mpnn_b200.heads.BNLinearMSE (csrc/head.cu) against the stock torch modules it wraps
(nn.BatchNorm1d -> nn.Linear -> nn.MSELoss, reference test_graph_norm.py:86-90), fp32 on the same device."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,C,T", [(256, 64, 12), (32, 38, 1), (7, 5, 3), (1024, 16, 1), (200, 100, 17)])
@pytest.mark.parametrize("mode", ["train", "eval", "no_affine", "untracked"])
def test_fused_head_matches_stock_modules(dev, B, C, T, mode):
    from mpnn_b200 import _lib
    from mpnn_b200.heads import BNLinearMSE
    assert _lib.load().mpnn_head_supported(B, C, T)
    torch.manual_seed(B + C + T)
    bn = nn.BatchNorm1d(C, affine=(mode != "no_affine"), track_running_stats=(mode != "untracked")).to(dev)
    lin = nn.Linear(C, T).to(dev)
    with torch.no_grad():
        if bn.affine:
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_()
        if bn.track_running_stats:
            bn.running_mean.normal_()
            bn.running_var.uniform_(0.5, 2.0)
    bn2, lin2 = copy.deepcopy(bn), copy.deepcopy(lin)
    fused = BNLinearMSE(bn2, lin2)
    if mode == "eval":
        bn.eval()
        fused.eval()
    x = (torch.randn(B, C, device=dev) * 2 + 0.5)
    t = torch.randn(B, T, device=dev)
    x1 = x.clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    y = lin(bn(x1))
    loss = torch.nn.functional.mse_loss(y, t)
    (loss * 1.7).backward()
    loss2 = fused(x2, t)
    (loss2 * 1.7).backward()
    assert rel_err(loss2, loss) <= 1e-5
    assert rel_err(fused.prediction, y) <= 1e-5
    assert rel_err(x2.grad, x1.grad) <= 1e-4
    for (k, p), (_, q) in zip(list(lin.named_parameters()) + list(bn.named_parameters()),
                              list(lin2.named_parameters()) + list(bn2.named_parameters())):
        assert rel_err(q.grad, p.grad) <= 1e-4, k
    for (k, p), (_, q) in zip(bn.named_buffers(), bn2.named_buffers()):
        assert rel_err(q.float(), p.float()) <= 1e-5, k


def test_fused_head_reproducible_and_refuses_unserved_shapes(dev):
    from mpnn_b200.heads import BNLinearMSE
    torch.manual_seed(0)
    bn, lin = nn.BatchNorm1d(64).to(dev), nn.Linear(64, 12).to(dev)
    fused = BNLinearMSE(bn, lin)
    x, t = torch.randn(256, 64, device=dev), torch.randn(256, 12, device=dev)
    outs = []
    for _ in range(2):
        xi = x.clone().requires_grad_(True)
        fused.zero_grad()
        fused(xi, t).backward()
        outs.append((xi.grad.clone(), lin.weight.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # larger than the kernel serves: an error, never a silent switch to another implementation
    xb, tb = torch.randn(8192, 64, device=dev), torch.randn(8192, 12, device=dev)
    with pytest.raises(RuntimeError):
        fused(xb, tb)
    with pytest.raises(RuntimeError):
        fused(x.cpu(), t.cpu())
