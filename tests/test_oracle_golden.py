"""CPU: the oracle restatement (oracle/mpnn_oracle.py) against the golden vectors frozen from the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from golden_util import Case, all_cases, rel_err, run_oracle

# the restatement calls the same aten ops in the same order, so agreement is ~1 ulp-level
TOL_OUT = 2e-6
TOL_GRAD = 2e-5


@pytest.mark.parametrize("name", all_cases())
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(2)
    case = Case(name)
    out, gin, gsd, buffers = run_oracle(case)
    assert out.shape == case.out["y"].shape
    assert rel_err(out, case.out["y"]) <= TOL_OUT
    for k, g in case.gin.items():
        assert rel_err(gin[k], g) <= TOL_GRAD, "grad of input %s" % k
    for k, g in case.gsd.items():
        assert rel_err(gsd[k], g) <= TOL_GRAD, "grad of parameter %s" % k
    for k, v in case.out.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_err(buffers[k if "." in k else k], v) <= TOL_OUT, k


def test_compaction_oracle_order():
    from oracle import mpnn_oracle as O
    from mpnn_b200 import synthetic
    b = synthetic.small_batch(B=4, n_lo=1, n_hi=7, afm_width=3, ef=3, seed=5, weighted_adj=True)
    bfm, adj = torch.from_numpy(b["bfm"]), torch.from_numpy(b["adj"])
    bfm[0, 0, 1] = 0  # an adj-only edge (bond row all zero) and a bfm-only edge must both be kept
    adj[1, 0, 1] = 0
    c = O.compact_edges(bfm, adj)
    N = adj.shape[1]
    E = int(c["row_ptr"][-1])
    assert E == len(c["dst"]) == len(c["src"])
    keys = c["dst"].long() * N + (c["src"].long() % N)
    assert bool((keys[1:] > keys[:-1]).all())  # row-major, strictly increasing
    for r in range(adj.shape[0] * N):
        seg = c["dst"][c["row_ptr"][r]:c["row_ptr"][r + 1]]
        assert bool((seg == r).all())


def test_layout_fixture_matches_collate():
    from golden_util import GOLD
    import os
    from mpnn_b200 import synthetic
    z = np.load(os.path.join(GOLD, "layout_collate.npz"))
    graphs = synthetic.make_graphs(5, ("uniform", 2, 9), 6, 5, nafm_width=3, seed=99)
    mine = synthetic.collate(graphs)
    for k in ("afm", "nafm", "bfm", "adj", "mask"):
        assert np.array_equal(z[k], mine[k])
