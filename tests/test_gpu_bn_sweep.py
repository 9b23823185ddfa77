"""MaskBatchNorm (reference models/mask_batch_norm.py:9-15, csrc/bn.cu) over a sweep of shapes -- single-row-block to
10^5 rows, 1 to 512 columns, 0/1 and weighted masks -- against the reference formula evaluated in fp64 by torch
autograd.  (A single-launch thread-block-cluster form of these kernels was measured and dropped: 13.5 / 11.0 us per
call against 10.9 / 8.5 us for the grid form at 7 424 x 16; its ~25 block-wide barriers of 1 024 threads cost more
than the second launch saves.  tools/bench_bn.py, tools/microbench/cluster_lat.cu.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def reference(x, mask, eps=1e-6):
    m = mask.reshape(-1, 1)
    y = x.reshape(-1, x.shape[-1])
    mean = y.sum(0) / m.sum()                      # unmasked sum (relies on zeros at the padding)
    c = (y - mean) * m
    var = c.pow(2).sum(0) / m.sum()
    return (c / torch.sqrt(var + eps)).view(x.shape)


@pytest.mark.parametrize("B,N,C", [(256, 29, 16), (4, 25, 19), (1, 63, 5), (128, 38, 32), (7, 9, 1), (2, 40, 300),
                                    (1000, 40, 32), (4096, 38, 8), (64, 30, 512)])
@pytest.mark.parametrize("weighted", [False, True])
def test_mask_bn_shape_sweep(B, N, C, weighted):
    from mpnn_b200.modules import MaskBatchNorm
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 31 + N * 7 + C)
    mask = (torch.rand(B, N, 1, generator=g) > 0.3).float()
    if weighted:
        mask = mask * torch.rand(B, N, 1, generator=g).add(0.5)
    x = (torch.randn(B, N, C, generator=g) * 1.5 + 0.7) * (mask > 0).float()
    cot = torch.randn(B, N, C, generator=g)
    xd = x.double().requires_grad_(True)
    ref = reference(xd, mask.double())
    (ref * cot.double()).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    out = MaskBatchNorm()(xg, mask.to(dev))
    (out * cot.to(dev)).sum().backward()
    assert rel_err(out.cpu(), ref) <= 2e-5
    assert rel_err(xg.grad.cpu(), xd.grad) <= 2e-4
    # padded rows are exactly zero, and reruns are bit-identical
    assert float(out[(mask == 0).to(dev).expand_as(out)].abs().max()) == 0.0
    xg2 = x.to(dev).requires_grad_(True)
    out2 = MaskBatchNorm()(xg2, mask.to(dev))
    (out2 * cot.to(dev)).sum().backward()
    assert torch.equal(out, out2) and torch.equal(xg.grad, xg2.grad)


def test_workspace_reuse_across_shapes_of_equal_size():
    """The batch norms reuse one persistent zeroed workspace per byte size.  Two shapes whose workspaces have the same
    size but a different internal layout must not disturb each other (round 2: the completion counter of one layout lay
    inside the partial sums of the other; the second batch norm then started from a non-zero counter and normalised with
    garbage).  Every pair of shapes with equal workspace size is run back to back, both orders."""
    from mpnn_b200 import _lib
    from mpnn_b200.modules import MaskBatchNorm
    lib = _lib.load()
    dev = torch.device("cuda:0")
    shapes = [(r, c) for r in (96, 128, 250, 256, 300, 512, 1000, 1024, 2048) for c in (8, 16, 24, 32, 40, 48, 64, 80, 128)]
    by_size = {}
    for r, c in shapes:
        by_size.setdefault(int(lib.mpnn_bn_workspace_bytes(r, c)), []).append((r, c))
    pairs = [(a, b) for v in by_size.values() for a in v for b in v if a != b]
    assert pairs, "no two shapes share a workspace size: widen the sweep"
    g = torch.Generator().manual_seed(0)

    def run(rows, C):
        mask = (torch.rand(1, rows, 1, generator=g) > 0.3).float()
        x = (torch.randn(1, rows, C, generator=g) * 3 + 1) * mask
        out = MaskBatchNorm()(x.to(dev), mask.to(dev))
        assert rel_err(out.cpu(), reference(x.double(), mask.double())) <= 2e-5, (rows, C)

    for a, b in pairs[:40]:
        run(*a)
        run(*b)
