"""CPU tests (no GPU): the C-ABI library builds, loads and exports every symbol the header declares; the
drop-in modules keep the reference's state_dict contract; host-side logic (lazy messages protocol, batch
sharding, gradient all-reduce over gloo with world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from golden_util import Case, ROOT, all_cases


@pytest.fixture(scope="session")
def lib_path():
    from mpnn_b200 import build
    return build.build_library()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mpnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(lib_path):
    from mpnn_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 35
    lib = ctypes.CDLL(lib_path)
    for s in syms:
        assert hasattr(lib, s), "libmpnn_b200.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and include/mpnn_b200.h disagree"
    # argument counts agree with the header
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mpnn_b200.h")).read(), flags=re.S)
    for s in syms:
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % s, text, flags=re.S)
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[s][1]), s


def test_version_and_error_calls_need_no_gpu(lib_path):
    from mpnn_b200 import _lib
    lib = _lib.load()
    assert lib.mpnn_version() == 100
    assert lib.mpnn_gemm_workspace_bytes(8, 8, 8) == 0
    assert lib.mpnn_edge_trunk_saved_floats(10, 7, 1, 49, 50, None, None) == 10 * 52 + 50 * 10 * 52
    assert lib.mpnn_edge_trunk_saved_floats(10, 7, 1, 50, 50, None, None) == -1  # 7 -> 49 != 50


def test_cpu_tensors_fail_loudly():
    from mpnn_b200 import modules as M
    net = M.EdgeNetwork(4, 2, 4)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 4), torch.zeros(1, 3, 3, 2))
    with pytest.raises(RuntimeError):
        M.GRUUpdate(4, 4)(torch.zeros(1, 3, 4), torch.zeros(1, 3, 4), torch.ones(1, 3, 1))
    with pytest.raises(RuntimeError):
        M.MaskBatchNorm()(torch.zeros(1, 3, 4), torch.ones(1, 3, 1))
    with pytest.raises(NotImplementedError):
        M.EdgeNetwork(4, 2, 4, activation_fn=torch.nn.Tanh())


@pytest.mark.parametrize("name", [n for n in all_cases() if not n.startswith("model_")])
def test_module_state_dict_contract(name):
    """same keys and shapes as the reference module's state_dict (incl. the 50 aliased tied-layer keys)"""
    from mpnn_b200 import modules as M
    case = Case(name)
    m = case.meta
    cls = m["cls"].replace("EdgeNetworkD", "EdgeNetwork")
    if cls in ("EdgeNetwork", "AttEdgeNetwork", "GGNNMsgPass", "BiLiniearEdgeNetwork"):
        mod = getattr(M, cls)(m["nf"], m["ef"], m["mf"])
    elif cls in ("AdjMsgAgg", "WAdjMsgAgg", "AttMsgAgg"):
        mod = getattr(M, cls)(1)
    elif cls == "GRUUpdate":
        mod = M.GRUUpdate(m["d"], m["d"])
    elif cls == "MaskBatchNorm":
        mod = M.MaskBatchNorm()
    elif cls == "MaskBatchNorm1d":
        mod = M.MaskBatchNorm1d(5)
    elif cls in ("GraphLevelOutput", "GraphLevelOutputAtoms"):
        mod = getattr(M, cls)(m["nf"], m["out"])
    elif cls == "LSTMCellHidden":
        mod = M.LSTMCellHidden(m["hd"], m["cd"])
    else:
        mod = M.Set2Vec(m["nf"], 99, time_steps=m["steps"])
    sd = mod.state_dict()
    assert sorted(sd) == sorted(case.sd)
    for k, v in case.sd.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    if cls in ("EdgeNetwork", "AttEdgeNetwork"):
        tied = [k for k in sd if re.match(r"edge_map\.\d+\.0\.weight", k)]
        assert len(tied) == 50 and len({sd[k].data_ptr() for k in tied}) == 1
        from oracle import mpnn_oracle as O
        growth, P, first, last = O.edge_map_layout(m["nf"], m["ef"], m["mf"])
        assert (mod.P, mod._tied_idx, mod._last_idx) == (P, first, last)
    mod.load_state_dict(case.sd, strict=True)


def test_lazy_messages_protocol():
    """unknown consumers get the HEAD tensor: attribute access, operators and torch functions materialise"""
    from mpnn_b200.modules import LazyMessages

    class FakeNet(object):
        calls = 0

        def _head_messages(self, afm, bfm, reuse):
            FakeNet.calls += 1
            return afm * 2.0

    afm = torch.arange(6.0).view(1, 3, 2)
    lazy = LazyMessages(FakeNet(), afm, None, False)
    assert FakeNet.calls == 0
    assert tuple(lazy.shape) == (1, 3, 2) and FakeNet.calls == 1
    assert torch.equal(lazy.view(-1, 2), (afm * 2).view(-1, 2))
    assert torch.equal(lazy + 1, afm * 2 + 1) and torch.equal(1 + lazy, afm * 2 + 1)
    assert torch.equal(torch.cat([lazy, afm], dim=-1), torch.cat([afm * 2, afm], dim=-1))
    assert torch.equal(afm.mul(lazy), afm * afm * 2)
    assert torch.equal(lazy.mul(afm).sum(dim=-2), (afm * afm * 2).sum(dim=-2))
    assert FakeNet.calls == 1


def test_reference_model_files_run_against_dropin_package():
    """`from mpnn_functions import *` / `from mask_batch_norm import ...` in the UNMODIFIED reference model files
    resolve to this package (build container only: /root/reference is not on the GPU box)."""
    ref_models = "/root/reference/models"
    if not os.path.isdir(ref_models):
        pytest.skip("reference tree not present")
    import subprocess
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from mpnn_b200 import dropin, modules
dropin.install()
import basic_model, normed_basic_model, lipo_basic_model, att_model, basic_graph_autoencoder
m = lipo_basic_model.BasicModel(22, 7, 22, 1, 38, message_opts={}, agg_opts={}, update_opts={}, readout_opts={})
assert type(m.mf) is modules.EdgeNetwork and type(m.bn) is modules.MaskBatchNorm1d
m.apply(lipo_basic_model.BasicModel.init_weights)
assert len(m.state_dict()) == 55 + 5 + 5 + 4 + 4, len(m.state_dict())
a = att_model.BasicModel(8, 3, 8, 1, 4, message_opts={}, agg_opts={}, update_opts={}, readout_opts={})
assert type(a.mfs[0]) is modules.AttEdgeNetwork and type(a.of) is modules.Set2Vec
n = normed_basic_model.BasicModel(16, 7, 16, 1, 64, message_opts={}, agg_opts={}, update_opts={}, readout_opts={})
assert type(n.bn) is modules.MaskBatchNorm
print("ok")
""" % (ROOT, ref_models)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_model_state_dict_contract():
    from mpnn_b200.dropin import reference_model as MessagePassingModel
    case = Case("model_lipo")
    mod = MessagePassingModel("lipo", case.meta["d"], case.meta["ef"], case.meta["d"], 1, case.meta["out"],
                              message_steps=3)
    assert sorted(mod.state_dict()) == sorted(case.sd)
    case = Case("model_normed_basic")
    mod = MessagePassingModel("normed", case.meta["d"], case.meta["ef"], case.meta["d"], 1, case.meta["out"],
                              message_steps=2)
    assert sorted(mod.state_dict()) == sorted(case.sd)


def test_synthetic_generator_is_deterministic_and_shaped():
    from mpnn_b200 import synthetic
    a = synthetic.make_batch("qm9", B=16)
    b = synthetic.make_batch("qm9", B=16)
    for k in ("afm", "bfm", "adj", "mask", "labels"):
        assert np.array_equal(a[k], b[k])
    assert a["afm"].shape[2] == 16 and a["bfm"].shape[3] == 7 and a["labels"].shape == (16, 12)
    assert np.array_equal(a["adj"], a["adj"].transpose(0, 2, 1))
    assert ((a["bfm"] != 0).any(-1) == (a["adj"] != 0)).all()         # bond row non-zero <=> bonded
    assert (a["adj"].sum(-1) <= 4).all()                              # degree <= 4
    n = a["mask"].sum()
    assert 1.6 * n <= a["n_edges"] <= 2.4 * n                          # e ~ 2.2 n


def test_shard_batch_repads_to_local_max():
    from mpnn_b200 import dist as D, synthetic
    b = synthetic.make_batch("qm9", B=16)
    parts = [D.shard_batch(b, r, 2) for r in range(2)]
    assert parts[0]["afm"].shape[0] + parts[1]["afm"].shape[0] == 16
    for p in parts:
        nmax = int(p["mask"].reshape(p["mask"].shape[0], -1).sum(1).max())
        assert p["afm"].shape[1] == nmax and p["bfm"].shape[1:3] == (nmax, nmax)
    assert sum(float(p["mask"].sum()) for p in parts) == float(b["mask"].sum())


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from mpnn_b200.dist import FlatGradAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    x = torch.arange(12.0).view(4, 3)[rank * 2:(rank + 1) * 2]
    lin(x).pow(2).sum().backward()
    FlatGradAllReduce(lin.parameters())(average=False)
    if rank == 0:
        torch.save([p.grad.clone() for p in lin.parameters()], out)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "g.pt")
    mp.spawn(_dp_worker, args=(2, 29611 + os.getpid() % 200, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    lin(torch.arange(12.0).view(4, 3)).pow(2).sum().backward()
    for g, p in zip(got, lin.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-6, atol=1e-6)


def _dp_uneven_worker(rank, world, port, out):
    import torch.distributed as dist
    from mpnn_b200.dist import FlatGradAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    a, b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    x = torch.arange(12.0).view(4, 3)[rank * 2:(rank + 1) * 2]
    # rank 1 never touches `b`: its gradients are None there, the bucket layout must not depend on that
    (a(x).pow(2).sum() + (b(x).sum() if rank == 0 else 0.0)).backward()
    FlatGradAllReduce(list(a.parameters()) + list(b.parameters()))(average=False)
    torch.save([p.grad.clone() for p in list(a.parameters()) + list(b.parameters())], out + str(rank))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_uneven_none_pattern(tmp_path):
    """ranks with different `grad is None` patterns reduce the same layout"""
    import torch.multiprocessing as mp
    out = str(tmp_path / "g.pt")
    mp.spawn(_dp_uneven_worker, args=(2, 29811 + os.getpid() % 200, out), nprocs=2, join=True)
    g0, g1 = torch.load(out + "0"), torch.load(out + "1")
    torch.manual_seed(0)
    a, b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    x = torch.arange(12.0).view(4, 3)
    (a(x).pow(2).sum() + b(x[:2]).sum()).backward()
    want = [p.grad for p in list(a.parameters()) + list(b.parameters())]
    for u, v, w in zip(g0, g1, want):
        assert torch.allclose(u, w, rtol=1e-6, atol=1e-6) and torch.allclose(v, w, rtol=1e-6, atol=1e-6)


def test_ragged_batch_matches_host_collate():
    """loader.RaggedBatch: the ragged form carries exactly the non-zero content of the padded
    batch that `collate_2d_graphs` (reference data_loader.py:50-70) builds, in row-major (b, i, j) order."""
    import numpy as np
    from mpnn_b200 import synthetic
    from mpnn_b200.loader import RaggedBatch
    b = synthetic.make_batch("qm9", B=9, return_graphs=True)
    rb = RaggedBatch.from_graphs(b["graphs"], b["labels"], pin=False)
    B, N = rb.B, rb.N
    assert (B, N) == b["afm"].shape[:2] and rb.Fa == b["afm"].shape[2] and rb.ef == b["bfm"].shape[3]
    afm = np.zeros_like(b["afm"]).reshape(B * N, -1)
    afm[rb.atom_row.numpy()] = rb.afm_cat.numpy()
    assert np.array_equal(afm.reshape(b["afm"].shape), b["afm"])
    mask = np.zeros(B * N, np.float32)
    mask[rb.atom_row.numpy()] = 1
    assert np.array_equal(mask.reshape(B, N, 1), b["mask"])
    keep = (b["bfm"] != 0).any(-1) | (b["adj"] != 0)
    bb, ii, jj = np.nonzero(keep)
    assert np.array_equal(rb.edge_dst.numpy(), (bb * N + ii).astype(np.int32))
    assert np.array_equal(rb.edge_j.numpy(), jj.astype(np.int32))
    assert np.array_equal(rb.edge_w.numpy(), b["adj"][bb, ii, jj])
    assert np.array_equal(rb.edge_x.numpy(), b["bfm"][bb, ii, jj])
    padded = sum(b[k].nbytes for k in ("afm", "bfm", "adj", "mask"))
    assert rb.nbytes() < padded / 4


def test_vendored_model_files_are_the_unmodified_reference():
    """tests/ref_models/models/*.py are byte-identical copies of the reference's model files (clients of the boundary,
    test fixtures only): checked against the committed SHA256SUMS always, and against /root/reference when present."""
    import hashlib
    d = os.path.join(ROOT, "tests", "ref_models")
    sums = dict(reversed(line.split()) for line in open(os.path.join(d, "SHA256SUMS")) if line.strip())
    files = sorted(f for f in os.listdir(os.path.join(d, "models")) if f.endswith(".py"))
    assert files == sorted(sums), (files, sorted(sums))
    for f in files:
        data = open(os.path.join(d, "models", f), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sums[f], f
        ref = os.path.join("/root/reference/models", f)
        if os.path.exists(ref):
            assert open(ref, "rb").read() == data, "%s differs from the reference" % f


def test_dropin_loads_every_model_file():
    """dropin.reference_models(): every vendored model file imports against this package and constructs (CPU)"""
    from mpnn_b200 import dropin, modules
    ns = dropin.reference_models()
    for variant, (fname, cls, _) in dropin.MODEL_FILES.items():
        assert hasattr(ns, fname), fname
        kw = {}
        if variant.startswith("normed_encoded"):
            kw = dict(atom_encoder=modules.AtomAutoEncoder().encoder, bond_encoder=modules.BondAutoEncoder().encoder)
        d, ef = (8, 2) if kw else (16, 7)
        m = dropin.reference_model(variant, d, ef, d, 1, 12, **kw)
        assert any(isinstance(c, modules.EdgeNetwork) for c in m.modules()), variant
        assert isinstance(m.uf, modules.GRUUpdate)
    w = ns.graph_model_wrapper.GraphWrapper(dropin.reference_model("basic", 16, 7, 16, 1, 12))
    assert "graph_model.mf.message_bias" in w.state_dict()
    assert isinstance(ns.graph_norm_wrapper.GraphWrapper(torch.nn.Identity(), 3).bn, modules.MaskBatchNorm1d)


def test_dropin_encoders_state_dict_contract():
    """mpnn_functions.encoders drop-in: same layers / state_dict keys as the reference's encoder classes
    (mpnn_functions/encoders/{atom,bond}_autoencoder.py:7-20, auto_encoder.py:7-19)"""
    from mpnn_b200 import dropin
    dropin.install()
    from mpnn_functions.encoders.atom_autoencoder import AtomAutoEncoder
    from mpnn_functions.encoders.bond_autoencoder import BondAutoEncoder
    from mpnn_functions.encoders.auto_encoder import Autoencoder
    a, b, c = AtomAutoEncoder(), BondAutoEncoder(), Autoencoder(12, 6, 3)
    assert {k: tuple(v.shape) for k, v in a.encoder.state_dict().items()} == {
        "0.weight": (15, 30), "2.weight": (8, 15), "2.bias": (8,)}
    assert {k: tuple(v.shape) for k, v in b.encoder.state_dict().items()} == {
        "0.weight": (4, 8), "2.weight": (2, 4), "2.bias": (2,)}
    assert sorted(a.state_dict()) == sorted(
        ["encoder.0.weight", "encoder.2.weight", "encoder.2.bias", "decoder.0.weight", "decoder.0.bias",
         "decoder.0.running_mean", "decoder.0.running_var", "decoder.0.num_batches_tracked", "decoder.1.weight",
         "decoder.1.bias", "decoder.3.weight", "decoder.3.bias"])
    assert sorted(c.state_dict()) == ["decoder.0.weight", "decoder.2.weight", "encoder.0.weight", "encoder.2.weight"]
    x = torch.randn(5, 8)
    assert b.encoder(x).shape == (5, 2) and b(x).shape == (5, 8)      # CPU tensors: the stock layers
    ref = "/root/reference/mpnn_functions/encoders"
    if os.path.isdir(ref):
        import importlib.util
        for name, ours in (("atom_autoencoder", a), ("bond_autoencoder", b)):
            spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(ref, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            theirs = [v for k, v in vars(mod).items() if k.endswith("AutoEncoder")][0]()
            assert {k: tuple(v.shape) for k, v in theirs.state_dict().items()} == \
                {k: tuple(v.shape) for k, v in ours.state_dict().items()}


def test_edge_list_slab_views_are_aligned_and_disjoint():
    """graph._slab_views (the one-allocation edge list of the pipelined preprocessing): consecutive views, each starting
    on a 256-byte boundary, none overlapping, every requested length honoured"""
    import torch
    from mpnn_b200 import graph
    sizes = [7425, 7425, 10794, 10794, 10794, 10794, 10794, 65 * 7, 4, 65, 10794, 10794, 1, 63, 64, 65]
    slab = torch.zeros(sum((n + 63) // 64 * 64 for n in sizes), dtype=torch.int32)
    views = graph._slab_views(slab, sizes)
    assert [v.numel() for v in views] == sizes
    end = 0
    for i, v in enumerate(views):
        off = v.storage_offset()
        assert off % 64 == 0 and off >= end, (i, off, end)       # 64 int32 = 256 bytes
        end = off + v.numel()
        v.fill_(i + 1)
    assert end <= slab.numel()
    for i, v in enumerate(views):
        assert bool((v == i + 1).all())
