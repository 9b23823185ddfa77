"""Synthetic molecule-shaped batches in the reference's padded layout (host side, numpy only).

The layout is the one ``collate_2d_graphs`` produces (reference
``pre_process/data_loader.py:50-70``): every graph is zero-padded to the largest
graph of the batch; ``mask[b, i, 0] = 1`` for real atoms; ``bfm[b, i, j, :] == 0``
wherever ``adj[b, i, j] == 0`` (``mol_graph/mol_graph.py:207-219``).  The graph
statistics follow SURVEY.md §8d: random tree with degree <= 4 plus ~10 % ring
closures, one-hot atom/bond feature blocks shaped like ``AtomFeatures`` /
``BondFeatures`` (``mol_graph/mol_graph.py:37-90``).

The generator is shared by the oracle, the tests and ``bench.py`` so both sides of
every parity check see identical bytes.
"""
import numpy as np

REFERENCE_SEED = 317  # reference test_lipo.py:74

# name -> (config_id, size law, feature widths); BASELINE.json `configs` order.
CONFIGS = {
    "lipo": dict(cid=1, B=32, size=("normal", 27, 8, 6, 64), afm=19, nafm=3, ef=7, d=22, T=6, out=38),
    "qm9": dict(cid=2, B=256, size=("uniform", 4, 29), afm=16, nafm=0, ef=7, d=16, T=3, out=64, targets=12),
    "zinc": dict(cid=3, B=128, size=("normal", 23, 5, 6, 38), afm=32, nafm=0, ef=8, d=32, T=3),
    "affinity": dict(cid=4, B=2048, size=("normal", 28, 7, 8, 50), afm=30, nafm=0, ef=8, d=8, T=3, out=16),
    "autoenc": dict(cid=5, B=512, size=("normal", 23, 5, 6, 38), afm=64, nafm=0, ef=8, d=64, T=3),
}


def _sizes(rs, law, B):
    if law[0] == "uniform":
        return rs.randint(law[1], law[2] + 1, size=B)
    _, mu, sd, lo, hi = law
    return np.clip(np.rint(rs.normal(mu, sd, size=B)), lo, hi).astype(np.int64)


def _skeleton(rs, n):
    """Random tree (degree <= 4) + round(0.1 n) ring closures at tree distance >= 4."""
    parent = np.full(n, -1, dtype=np.int64)
    deg = np.zeros(n, dtype=np.int64)
    depth = np.zeros(n, dtype=np.int64)
    edges = []
    for i in range(1, n):
        while True:
            p = int(rs.randint(0, i))
            if deg[p] < 4:
                break
        parent[i] = p
        depth[i] = depth[p] + 1
        deg[p] += 1
        deg[i] += 1
        edges.append((p, i))

    def tree_dist(a, b):
        d = 0
        while a != b:
            if depth[a] < depth[b]:
                a, b = b, a
            a = parent[a]
            d += 1
        return d

    want = int(round(0.1 * n))
    tries = 0
    have = set(edges)
    while want > 0 and tries < 20 * n:
        tries += 1
        a, b = int(rs.randint(0, n)), int(rs.randint(0, n))
        if a == b or deg[a] >= 4 or deg[b] >= 4:
            continue
        if (min(a, b), max(a, b)) in have or tree_dist(a, b) < 4:
            continue
        have.add((min(a, b), max(a, b)))
        edges.append((min(a, b), max(a, b)))
        deg[a] += 1
        deg[b] += 1
        want -= 1
    return np.asarray(edges, dtype=np.int64).reshape(-1, 2)


def _one_hot_block(rs, n, width, p=None):
    idx = rs.choice(width, size=n, p=p)
    out = np.zeros((n, width), dtype=np.float32)
    out[np.arange(n), idx] = 1.0
    return out


def make_graphs(B, size_law, afm_width, ef, nafm_width=0, seed=REFERENCE_SEED):
    """List of ragged graphs: dicts with afm [n,Fa], nafm [n,Fn], bfm [n,n,ef], adj [n,n]."""
    rs = np.random.RandomState(seed)
    sizes = _sizes(rs, size_law, B)
    hyb = max(2, min(8, afm_width // 4))
    elem = afm_width - hyb - 2
    assert elem >= 1, "afm width too small for the AtomFeatures-shaped blocks"
    n_types = ef - 3
    assert n_types >= 1
    p_types = np.array([0.70, 0.15, 0.12, 0.03] + [0.0] * max(0, n_types - 4), dtype=np.float64)[:n_types]
    p_types = p_types / p_types.sum()
    graphs = []
    for n in sizes:
        n = int(n)
        edges = _skeleton(rs, n)
        p_elem = np.ones(elem) / elem
        afm = np.concatenate([_one_hot_block(rs, n, elem, p_elem), _one_hot_block(rs, n, hyb),
                              (rs.rand(n, 2) < 0.4).astype(np.float32)], axis=1)
        adj = np.zeros((n, n), dtype=np.float32)
        bfm = np.zeros((n, n, ef), dtype=np.float32)
        if len(edges):
            feat = np.concatenate([_one_hot_block(rs, len(edges), n_types, p_types),
                                   (rs.rand(len(edges), 3) < 0.3).astype(np.float32)], axis=1)
            a, b = edges[:, 0], edges[:, 1]
            adj[a, b] = 1.0
            adj[b, a] = 1.0
            bfm[a, b] = feat
            bfm[b, a] = feat
        g = dict(afm=afm, bfm=bfm, adj=adj)
        if nafm_width:
            g["nafm"] = rs.rand(n, nafm_width).astype(np.float32)  # min-max scaled numeric columns
        graphs.append(g)
    return graphs


def collate(graphs, labels=None):
    """Zero-pad to the largest graph, exactly like reference data_loader.py:50-70."""
    B = len(graphs)
    N = max(g["afm"].shape[0] for g in graphs)
    Fa = graphs[0]["afm"].shape[1]
    ef = graphs[0]["bfm"].shape[2]
    out = dict(
        afm=np.zeros((B, N, Fa), np.float32), bfm=np.zeros((B, N, N, ef), np.float32),
        adj=np.zeros((B, N, N), np.float32), mask=np.zeros((B, N, 1), np.float32))
    if "nafm" in graphs[0]:
        out["nafm"] = np.zeros((B, N, graphs[0]["nafm"].shape[1]), np.float32)
    for b, g in enumerate(graphs):
        n = g["afm"].shape[0]
        out["afm"][b, :n] = g["afm"]
        out["bfm"][b, :n, :n] = g["bfm"]
        out["adj"][b, :n, :n] = g["adj"]
        out["mask"][b, :n] = 1.0
        if "nafm" in g:
            out["nafm"][b, :n] = g["nafm"]
    if labels is not None:
        out["labels"] = labels
    return out


def make_batch(name, B=None, seed_offset=0, d=None, return_graphs=False):
    """Padded batch dict for one of the BASELINE.json configs (numpy, float32)."""
    cfg = dict(CONFIGS[name])
    if B is not None:
        cfg["B"] = B
    afm_w = cfg["afm"] if d is None else d
    graphs = make_graphs(cfg["B"], cfg["size"], afm_w, cfg["ef"], cfg["nafm"],
                         seed=REFERENCE_SEED + cfg["cid"] + 1000 * seed_offset)
    rs = np.random.RandomState(REFERENCE_SEED + cfg["cid"] + 7919 + 1000 * seed_offset)
    n_t = cfg.get("targets", 1)
    labels = rs.normal(size=(cfg["B"], n_t)).astype(np.float32)
    batch = collate(graphs, labels)
    batch["n_atoms"] = int(batch["mask"].sum())
    batch["n_edges"] = int((batch["adj"] != 0).sum())
    if return_graphs:
        batch["graphs"] = graphs
    return batch


def small_batch(B=3, n_lo=2, n_hi=7, afm_width=6, ef=4, nafm_width=0, seed=0, weighted_adj=False):
    """Tiny ragged batch for unit tests (may contain single-atom / edgeless graphs)."""
    rs = np.random.RandomState(seed)
    graphs = []
    for _ in range(B):
        n = int(rs.randint(n_lo, n_hi + 1))
        edges = _skeleton(rs, n) if n > 1 else np.zeros((0, 2), np.int64)
        afm = rs.normal(size=(n, afm_width)).astype(np.float32)
        adj = np.zeros((n, n), np.float32)
        bfm = np.zeros((n, n, ef), np.float32)
        for a, b in edges:
            w = float(rs.uniform(0.5, 2.0)) if weighted_adj else 1.0
            adj[a, b] = adj[b, a] = w
            f = rs.normal(size=ef).astype(np.float32)
            bfm[a, b] = bfm[b, a] = f
        g = dict(afm=afm, bfm=bfm, adj=adj)
        if nafm_width:
            g["nafm"] = rs.rand(n, nafm_width).astype(np.float32)
        graphs.append(g)
    return collate(graphs, rs.normal(size=(B, 1)).astype(np.float32))
