"""GPU tests of the tensor-core typed path (csrc/tc_message.cu: tcgen05.mma kind::tf32, accumulator in TMEM) for
feature widths 33..256: the grouped edge GEMM + segmented sums against an fp64 torch restatement of
edge_network.py:52 + adjacent_message_agg.py:18 on the same edge list, and the modules against the fp32 per-edge
contraction kernels (csrc/message.cu), which the golden-vector tests pin to the reference.

Tolerance: operands are read as TF32 (10-bit mantissa, fp32 accumulate) -> max-abs error <= 3e-3 x max-abs(ref)
(SURVEY.md 8c allows 2e-2 for tensor-core inputs); everything else on the path stays fp32."""
import numpy as np
import pytest
import torch

from golden_util import rel_err
from test_gpu_typed import _net, _run

pytestmark = pytest.mark.gpu

TF32_TOL = 3e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpnn_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _categorical_batch(B, ef, seed=0):
    """ZINC-shaped graphs with categorical bond rows (few distinct rows), bond width remapped to `ef`"""
    from mpnn_b200 import synthetic
    b = synthetic.make_batch("autoenc", B=B, seed_offset=seed)
    bfm = b["bfm"]
    if ef != bfm.shape[-1]:
        R = np.random.RandomState(11).normal(size=(bfm.shape[-1], ef)).astype(np.float32)
        bfm = (bfm @ R).astype(np.float32)
    return bfm, b["adj"]


@pytest.mark.parametrize("dims", [(64, 64, 40), (96, 128, 24), (256, 256, 12), (36, 64, 8)])
@pytest.mark.parametrize("weighted", [False, True])
def test_tc_kernels_against_fp64(dev, dims, weighted):
    from mpnn_b200 import graph
    from mpnn_b200.functional import TypedMessageTCFn, tc_dp
    nf, mf, B = dims
    bfm, adj = _categorical_batch(B, 8, seed=nf)
    g = torch.Generator().manual_seed(nf + mf)
    if weighted:
        w = torch.rand(adj.shape, generator=g).numpy() + 0.5
        adj = adj * np.maximum(w, np.swapaxes(w, 1, 2))
    bfm_t, adj_t = torch.from_numpy(bfm).to(dev), torch.from_numpy(adj.astype(np.float32)).to(dev)
    graph.clear_cache()
    el = graph.compact_edges(bfm_t, adj_t)
    ti = el.typed()
    assert ti.type_ptr is not None and ti.Ucap < 200
    DP = tc_dp(nf, mf)
    assert DP in (64, 128, 256)
    table = torch.zeros(ti.Ucap + 1, DP, DP)
    table[:, :nf, :mf] = torch.randn(ti.Ucap + 1, nf, mf, generator=g)
    table = table.to(dev).requires_grad_(True)
    tableT = table.detach().transpose(1, 2).contiguous()
    H = torch.randn(el.n_rows, nf, generator=g).to(dev).requires_grad_(True)
    cot = torch.randn(el.n_rows, mf, generator=g).to(dev)
    M = TypedMessageTCFn.apply(H, table, tableT, el, True, nf, mf)
    (M * cot).sum().backward()

    H64 = H.detach().double().requires_grad_(True)
    T64 = table.detach().double().requires_grad_(True)
    src, dst, uid = el.edge_src.long(), el.edge_dst.long(), ti.uid.long()
    # msg_e[k] = sum_l T[u_e][l][k] h[src_e][l]
    msg = torch.einsum("elk,el->ek", T64[uid][:, :nf, :mf], H64[src]) * el.edge_w.double().unsqueeze(1)
    ref = torch.zeros(el.n_rows, mf, dtype=torch.float64, device=dev).index_add(0, dst, msg)
    (ref * cot.double()).sum().backward()
    assert rel_err(M.detach().cpu(), ref.detach().float().cpu()) <= TF32_TOL
    assert rel_err(H.grad.cpu(), H64.grad.float().cpu()) <= TF32_TOL
    assert rel_err(table.grad.cpu(), T64.grad.float().cpu()) <= TF32_TOL
    # padding of the table gradient and rows without edges are exact zeros
    assert float(table.grad[:, nf:, :].abs().max() if nf < DP else 0.0) == 0.0
    assert float(table.grad[ti.zero_type].abs().max()) == 0.0
    deg = (el.row_ptr[1:] - el.row_ptr[:-1]).long()
    assert float(M.detach()[deg == 0].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(64, 8, 64, 6), (40, 8, 72, 5), (128, 12, 128, 3)])
@pytest.mark.parametrize("form", ["agg", "head"])
def test_tc_modules_match_contraction_path(dev, shape, form):
    """EdgeNetwork (+AdjMsgAgg) through the tensor-core typed path == the fp32 per-edge contraction kernels"""
    nf, ef, mf, B = shape
    bfm, adj = _categorical_batch(B, ef, seed=3)
    g = torch.Generator().manual_seed(5)
    N = adj.shape[1]
    mask = torch.from_numpy((adj.sum(-1) > 0).astype(np.float32))
    afm = (torch.randn(B, N, nf, generator=g) * mask.unsqueeze(-1)).to(dev)
    bfm_t, adj_t = torch.from_numpy(bfm).to(dev), torch.from_numpy(adj).to(dev)
    net = _net(nf, ef, mf, dev, seed=ef)
    cot = torch.randn(B, N, mf, generator=g).to(dev)
    o1, ga1, gp1 = _run(net, afm, bfm_t, adj_t, form, True, cot)
    o0, ga0, gp0 = _run(net, afm, bfm_t, adj_t, form, False, cot)
    assert rel_err(o1.cpu(), o0.cpu()) <= TF32_TOL
    assert rel_err(ga1.cpu(), ga0.cpu()) <= TF32_TOL
    scale = max(float(v.abs().max()) for v in gp0.values())
    for k in gp0:
        diff = float((gp1[k] - gp0[k]).abs().max())
        assert diff <= 2 * TF32_TOL * float(gp0[k].abs().max()) + 1e-5 * scale, k


def test_tc_path_bit_reproducible(dev):
    nf = mf = 64
    bfm, adj = _categorical_batch(48, 8, seed=7)
    g = torch.Generator().manual_seed(1)
    afm = torch.randn(adj.shape[0], adj.shape[1], nf, generator=g).to(dev)
    bfm_t, adj_t = torch.from_numpy(bfm).to(dev), torch.from_numpy(adj).to(dev)
    net = _net(nf, 8, mf, dev, seed=2)
    cot = torch.randn(adj.shape[0], adj.shape[1], mf, generator=g).to(dev)
    r1 = _run(net, afm, bfm_t, adj_t, "agg", True, cot)
    r2 = _run(net, afm, bfm_t, adj_t, "agg", True, cot)
    assert torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1])
    for k in r1[2]:
        assert torch.equal(r1[2][k], r2[2][k]), k


@pytest.mark.parametrize("d", [64, 36, 100, 128, 256])
def test_gru_gate_gemms_on_tensor_cores(dev, d):
    """GRUUpdate at widths 33..256: the gate products run on the tcgen05 dense-GEMM mode (TF32 operands); outputs and
    every gradient against the fp32 CPU oracle (gru_update.py:26-35,66-68)."""
    from mpnn_b200 import modules as M
    from oracle import mpnn_oracle as O
    B, N = 7, 45          # 315 rows: two full tiles and a ragged one
    g = torch.Generator().manual_seed(d)
    mask = (torch.rand(B, N, 1, generator=g) > 0.2).float()
    m = torch.randn(B, N, d, generator=g)
    h = torch.randn(B, N, d, generator=g) * mask
    cot = torch.randn(B, N, d, generator=g)
    torch.manual_seed(d)
    mod = M.GRUUpdate(d, d)
    with torch.no_grad():
        mod.gru_cell.bias_ih.normal_(std=0.3)
        mod.gru_cell.bias_hh.normal_(std=0.3)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    mod = mod.to(dev)
    md, hd = m.clone().to(dev).requires_grad_(True), h.clone().to(dev).requires_grad_(True)
    out = mod(md, hd, mask.to(dev))
    (out * cot.to(dev)).sum().backward()
    m0, h0 = m.clone().requires_grad_(True), h.clone().requires_grad_(True)
    ref = O.gru_update(m0, h0, mask, sd, "")
    (ref * cot).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= TF32_TOL
    assert rel_err(md.grad.cpu(), m0.grad) <= TF32_TOL
    assert rel_err(hd.grad.cpu(), h0.grad) <= TF32_TOL
    for k, p in mod.named_parameters():
        assert rel_err(p.grad.cpu(), sd[k].grad) <= TF32_TOL, k
    assert float((out.detach().cpu() * (1 - mask)).abs().max()) == 0.0


@pytest.mark.parametrize("dims", [(64, 128), (100, 72), (256, 512)])
@pytest.mark.parametrize("masked", [True, False])
def test_readout_projections_on_tensor_cores(dev, dims, masked):
    """GraphLevelOutput with 2*nf / output_dim in the tensor-core range (graph_level_output.py:30-47): the i / j
    projections, their data and weight gradients run on the tcgen05 dense-GEMM mode (K and N cut into <= 256 blocks)."""
    from mpnn_b200 import modules as M
    from oracle import mpnn_oracle as O
    nf, out_dim = dims
    B, N = 5, 33
    g = torch.Generator().manual_seed(nf)
    mask = (torch.rand(B, N, 1, generator=g) > 0.25).float()
    x = torch.randn(B, N, 2 * nf, generator=g) * 0.5
    cot = torch.randn(B, out_dim, generator=g)
    torch.manual_seed(nf)
    mod = M.GraphLevelOutput(nf, out_dim)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    mod = mod.to(dev)
    xd = x.clone().to(dev).requires_grad_(True)
    out = mod(xd, mask=mask.to(dev) if masked else None)
    (out * cot.to(dev)).sum().backward()
    x0 = x.clone().requires_grad_(True)
    ref = O.graph_level_output(x0, mask if masked else None, sd, "")
    (ref * cot).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= TF32_TOL
    assert rel_err(xd.grad.cpu(), x0.grad) <= TF32_TOL
    for k, p in mod.named_parameters():
        assert rel_err(p.grad.cpu(), sd[k].grad) <= TF32_TOL, k


def test_wide_trunk_against_oracle(dev):
    """P = 256 trunk (ef = 16 -> one growth layer 16 -> 256, then 50 tied 256 x 256 layers: the layer plan of the
    reference's hidden >= 128 models, edge_network.py:15-21) on the distinct bond rows: skinny fp32 GEMMs forward and
    for the data gradient, the tied-weight gradient as one stacked X^T D GEMM on the tensor cores; against the CPU
    oracle (per-pair messages + AdjMsgAgg)."""
    from mpnn_b200 import graph, modules as M
    from oracle import mpnn_oracle as O
    from golden_util import leaf_sd
    nf = mf = 40
    bfm, adj = _categorical_batch(3, 16, seed=5)
    g = torch.Generator().manual_seed(8)
    N = adj.shape[1]
    mask = torch.from_numpy((adj.sum(-1) > 0).astype(np.float32))
    afm = torch.randn(3, N, nf, generator=g) * mask.unsqueeze(-1)
    bfm_t, adj_t = torch.from_numpy(bfm), torch.from_numpy(adj)
    net = _net(nf, 16, mf, dev, seed=6)
    assert net.P == 256
    sd = leaf_sd({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    graph.clear_cache()
    a = afm.clone().to(dev).requires_grad_(True)
    out = M.AdjMsgAgg(1)(net(a, bfm_t.to(dev)), adj_t.to(dev))
    cot = torch.randn(out.shape, generator=g)
    (out * cot.to(dev)).sum().backward()
    a0 = afm.clone().requires_grad_(True)
    ref = (O.edge_network_pairs(a0, bfm_t, sd, "", nf) * adj_t.unsqueeze(-1)).sum(-2)
    (ref * cot).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) <= TF32_TOL
    assert rel_err(a.grad.cpu(), a0.grad) <= TF32_TOL
    params = dict(net.named_parameters())
    gscale = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    for k, v in sd.items():
        if v.grad is None or k == "message_bias" or k not in params:
            continue
        diff = float((params[k].grad.cpu() - v.grad).abs().max())
        assert diff <= 2 * TF32_TOL * float(v.grad.abs().max()) + 1e-5 * gscale, k


@pytest.mark.parametrize("cfg", [(2048, 64), (512, 128)])
def test_config5_full_size_properties(dev, cfg):
    """BASELINE config 5 (basic_graph_autoencoder.encode, ZINC-shaped) at full width on the tensor-core path --
    hidden 64 (P = 64) at B = 2048 and hidden 128 (P = 4096: the 84 M-parameter trunk on the distinct rows) at B = 512:
    size-independent properties.  Run-to-run bit-identical (no atomics anywhere), graphs independent (permuting the
    batch permutes the outputs, within the TF32 summation-order noise), padded atoms contribute nothing to the readout."""
    from mpnn_b200 import synthetic, graph
    from mpnn_b200.dropin import reference_model as MessagePassingModel, kaiming_init
    B, d = cfg
    torch.manual_seed(317)
    batch = synthetic.make_batch("autoenc", B=B, d=d)
    t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    mod = MessagePassingModel("autoencoder", d, 8, d, 1, 2 * d, message_steps=3)
    mod.apply(kaiming_init)
    mod = mod.to(dev)
    assert mod.mf.P == (64 if d == 64 else 4096)

    def run(tt):
        graph.clear_cache()
        a = tt["afm"].clone().requires_grad_(True)
        o = mod(a, tt["bfm"], tt["adj"], tt["mask"])
        mod.zero_grad(set_to_none=True)
        o.pow(2).mean().backward()
        torch.cuda.synchronize()
        return o.detach(), a.grad.detach(), [p.grad.clone() for p in mod.parameters() if p.grad is not None]

    o1, g1, p1 = run(t)
    o2, g2, p2 = run(t)
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()
    assert torch.equal(o1, o2) and torch.equal(g1, g2)
    for x, y in zip(p1, p2):
        assert torch.equal(x, y)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).to(dev)
    tp = {k: v[perm].contiguous() for k, v in t.items()}
    o3, g3, _ = run(tp)
    assert rel_err(o3.cpu(), o1[perm].cpu()) <= TF32_TOL
    assert rel_err(g3.cpu(), g1[perm].cpu()) <= TF32_TOL
    # padded atoms: no gradient reaches their (zero) input features through the masked update and readout
    pad = (1 - t["mask"]).bool().expand_as(g1)
    assert float(g1[pad].abs().max()) <= 1e-6 * float(g1.abs().max())


@pytest.mark.parametrize("variant,d,T", [("autoencoder", 64, 3), ("basic", 40, 3), ("normed", 64, 2)])
def test_wide_chain_on_real_rows_equals_padded_evaluation(dev, variant, d, T, monkeypatch):
    """modules._wide_chain (tensor-core widths: the step loop on the REAL rows only, node tensors gathered once and
    scattered back once) == the same modules evaluated link by link on the padded rows: outputs, input gradients and every
    parameter gradient (same kernels, same TF32 products per row; only the batch-norm reductions see the rows in another
    tiling)"""
    from mpnn_b200 import graph, modules as M, synthetic
    from mpnn_b200.dropin import reference_model, kaiming_init
    b = synthetic.make_batch("zinc", B=24, d=d)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    torch.manual_seed(11)
    mod = reference_model(variant, d, 8, d, 1, 2 * d, message_steps=T)
    mod.apply(kaiming_init)
    with torch.no_grad():       # tame the 50-layer trunk: messages of O(1), so TF32 round-off does not dominate the check
        for net in ([mod.mf] if hasattr(mod, "mf") else mod.mfs):
            net.edge_map[net._last_idx].weight.mul_(0.05)
    mod = mod.to(dev).train()
    res = []
    for compact in (True, False):
        monkeypatch.setattr(M, "WIDE_COMPACT", compact)
        graph.clear_cache()
        mod.zero_grad()
        a = t["afm"].clone().requires_grad_(True)
        out = mod(a, t["bfm"], t["adj"], t["mask"])
        cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).to(dev)
        (out * cot).sum().backward()
        res.append((out.detach(), a.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}))
    (o1, a1, g1), (o0, a0, g0) = res
    assert rel_err(o1.cpu(), o0.cpu()) <= 2e-5
    assert rel_err(a1.cpu(), a0.cpu()) <= 2e-4
    assert float((a1 * (1 - t["mask"])).abs().max()) == 0.0 or variant != "autoencoder"
    scale = max(float(v.abs().max()) for v in g0.values())
    for k in g0:
        assert float((g1[k] - g0[k]).abs().max()) <= 2e-4 * float(g0[k].abs().max()) + 1e-6 * scale, k


@pytest.mark.parametrize("d,B", [(64, 40), (40, 12), (128, 16), (256, 6)])
def test_aggregation_inside_gru_kernel_is_bit_identical(dev, d, B, monkeypatch):
    """north star: aggregation fused with the GRU gates.  The CSR sum of the per-edge messages done by the fused GRU
    kernel's operand producer (mpnn_gru_fwd_agg) == mpnn_segment_sum followed by the same kernel: same summation order,
    so outputs, messages saved for the backward and every gradient are bit-identical"""
    from mpnn_b200 import graph, modules as M, synthetic
    from mpnn_b200.dropin import reference_model, kaiming_init
    b = synthetic.make_batch("zinc", B=B, d=d)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
    torch.manual_seed(5)
    mod = reference_model("normed", d, 8, d, 1, 2 * d, message_steps=2)
    mod.apply(kaiming_init)
    with torch.no_grad():
        for net in mod.mfs:
            net.edge_map[net._last_idx].weight.mul_(0.05)
    mod = mod.to(dev).train()
    res = []
    for fused in (True, False):
        monkeypatch.setattr(M, "AGG_IN_GRU", fused)
        graph.clear_cache()
        mod.zero_grad()
        a = t["afm"].clone().requires_grad_(True)
        out = mod(a, t["bfm"], t["adj"], t["mask"])
        cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).to(dev)
        (out * cot).sum().backward()
        res.append((out.detach(), a.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}))
    (o1, a1, g1), (o0, a0, g0) = res
    assert torch.equal(o1, o0) and torch.equal(a1, a0)
    for k in g0:
        assert torch.equal(g1[k], g0[k]), k


@pytest.mark.parametrize("rows,d", [(1000, 64), (333, 40), (5000, 64), (77, 36), (5, 64), (130, 44), (40000, 48)])
def test_gru_backward_one_pass_equals_separate_pointwise(dev, rows, d):
    """widths 33..64: the pointwise GRU backward inside the weight-gradient kernel's producers (mpnn_tc_gru_param_point)
    and of the data-gradient kernel (mpnn_tc_gru_data_grad) against the separate pointwise launch + grouped product:
    same gate gradients and the same MMA order for dW (bit-identical); dm / dh accumulate the K blocks in another order
    and the bias gradients are summed over another partition of the rows"""
    from mpnn_b200 import _lib, functional as Fn
    lib = _lib.load()
    g = torch.Generator().manual_seed(rows + d)
    m = torch.randn(rows, d, generator=g).to(dev)
    h = torch.randn(rows, d, generator=g).to(dev)
    mask = (torch.rand(rows, generator=g) > 0.25).float().to(dev)
    W = [(torch.randn(d, 3 * d, generator=g) * 0.2).to(dev).requires_grad_(True) for _ in range(2)]
    bb = [(torch.randn(3 * d, generator=g) * 0.1).to(dev).requires_grad_(True) for _ in range(2)]
    cot = torch.randn(rows, d, generator=g).to(dev)
    res = []
    for one in (1, 0):
        prev = lib.mpnn_gru_bwd_one_pass(one)
        try:
            mm, hh = m.clone().requires_grad_(True), h.clone().requires_grad_(True)
            out = Fn.GRUFn.apply(mm, hh, mask, W[0], W[1], bb[0], bb[1], None)
            grads = torch.autograd.grad(out, [mm, hh] + W + bb, cot)
            torch.cuda.synchronize()
            res.append([t.clone() for t in grads])
        finally:
            lib.mpnn_gru_bwd_one_pass(prev)
    a, b = res
    for i in (2, 3):
        assert torch.equal(a[i], b[i]), i
    for i in (0, 1, 4, 5):
        assert rel_err(a[i], b[i]) <= 1e-5, i
