"""CPU oracle: a functional restatement of hochshi/mpnn's message-passing path.

TEST INFRASTRUCTURE ONLY -- not part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file, and only as the checker / the reported CPU baseline.
The shipped path (``mpnn_b200``) never imports it and has no CPU fallback.

Parity status: the reference has no tests, golden vectors or published numbers
 so parity is pinned the other way round: ``oracle/make_golden.py``
imports the UNMODIFIED reference modules from /root/reference (``oracle/ref_loader.py``),
runs them on seeded inputs and freezes inputs, weights, outputs and gradients in
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this restatement
against those vectors (CPU, every round).  The arithmetic is the same torch ``aten``
ops the reference calls, in the same order, so the restatement is also what
``bench.py`` times as the reference's CPU path (kind "port").

Every function takes the reference's ``state_dict`` entries (same key names) in a
plain dict ``sd`` with a key ``prefix`` and cites the reference lines it follows.
"""
import math

import torch

_BIG_NEGATIVE = -1e8  # graph_level_output.py:4, set2vec.py:10


# ----------------------------------------------------------------------------------------------
# edge network (message function)
# ----------------------------------------------------------------------------------------------
def edge_map_layout(nf, ef, mf):
    """Layer plan of EdgeNetwork.edge_map (edge_network.py:14-21).

    Returns (growth, P, first_tied, last): growth = [(key_index, in, out)], tied layers live at
    sequential indices first_tied .. first_tied+49 (all aliases of one Linear), last Linear at `last`.
    """
    growth = []
    width = ef
    idx = 0
    while width ** 2 < nf * mf:
        growth.append((idx, width, width ** 2))
        width = width ** 2
        idx += 2  # Linear + activation entries
    return growth, width, idx, idx + 50


def edge_map(rows, sd, prefix, nf, ef, mf, n_tied=50):
    """edge_network.py:15-21 applied to rows [R, ef] -> [R, mf*nf]; ReLU activation (:13)."""
    growth, P, first_tied, last = edge_map_layout(nf, ef, mf)
    a = rows
    for idx, _, _ in growth:
        a = torch.relu(torch.nn.functional.linear(a, sd[prefix + "edge_map.%d.weight" % idx],
                                                  sd[prefix + "edge_map.%d.bias" % idx]))
    w_t = sd[prefix + "edge_map.%d.0.weight" % first_tied]
    for _ in range(n_tied):
        a = torch.relu(torch.nn.functional.linear(a, w_t))
    return torch.nn.functional.linear(a, sd[prefix + "edge_map.%d.weight" % last],
                                      sd[prefix + "edge_map.%d.bias" % last])


def edge_embed_head(afm, bfm, sd, prefix, mf):
    """EdgeNetwork._precompute_edge_embed at HEAD (edge_network.py:30-39): [B, N*mf, N*nf]."""
    B, N, nf = afm.shape
    ef = bfm.shape[-1]
    emb = edge_map(bfm.reshape(-1, ef), sd, prefix, nf, ef, mf).view(B, N, N, mf, nf)   # :36-37
    return emb.permute(0, 1, 3, 2, 4).contiguous().view(-1, N * mf, N * nf)              # :38


def edge_network_head(afm, bfm, sd, prefix, mf, emb=None):
    """EdgeNetwork.forward at HEAD (edge_network.py:42-51): bmm with the (cached) embedding, + bias."""
    B, N, nf = afm.shape
    if emb is None:                                                                       # :43-44
        emb = edge_embed_head(afm, bfm, sd, prefix, mf)
    messages = emb.bmm(afm.reshape(B, N * nf, 1)).view(B, N, mf)                          # :50
    return messages + sd[prefix + "message_bias"]                                         # :51


def edge_embed_pairs(afm, bfm, sd, prefix, mf):
    """Commented reference line edge_network.py:40: [B,N,N,mf,nf]."""
    nf = afm.shape[-1]
    ef = bfm.shape[-1]
    return edge_map(bfm, sd, prefix, nf, ef, mf).view(bfm.shape[:3] + (mf, nf))


def edge_network_pairs(afm, bfm, sd, prefix, mf, emb=None):
    """Documented per-pair form (commented reference lines edge_network.py:40 and :52): [B,N,N,mf], no bias."""
    if emb is None:
        emb = edge_embed_pairs(afm, bfm, sd, prefix, mf)                                  # :40
    return emb.matmul(afm.unsqueeze(1).unsqueeze(-1)).squeeze(-1)                          # :52


def att_edge_network_pairs(afm, bfm, sd, prefix, mf):
    """AttEdgeNetwork.forward (att_edge_network.py:13-31) over the line-40 edge embedding."""
    nf = afm.shape[-1]
    ef = bfm.shape[-1]
    emb = edge_map(bfm, sd, prefix, nf, ef, mf).view(bfm.shape[:3] + (mf, nf))
    cat = torch.cat((afm.unsqueeze(-2).expand(-1, -1, afm.shape[1], -1), bfm), dim=-1)    # :18
    attn_w = torch.softmax(torch.nn.functional.linear(cat, sd[prefix + "attn.weight"],
                                                      sd[prefix + "attn.bias"]), dim=-1)  # :21, default act :11
    attn_app = attn_w.mul(afm.unsqueeze(1)).unsqueeze(-1)                                  # :26
    return emb.matmul(attn_app).squeeze(-1)                                                # :31


def bilinear_edge_network(afm, bfm, nf):
    """BiLiniearEdgeNetwork.forward (bilinear_edge_network.py:25-37): parameter-free, needs ef == nf**3;
    out[b,i,j,p] = sum_{a,q} afm[b,j,a] * bfm[b,i,j].view(nf,nf,nf)[a,p,q] * afm[b,i,q]   -> [B,N,N,nf]."""
    ees = bfm.shape[:3] + (nf, -1)
    return afm.unsqueeze(1).unsqueeze(-2).matmul(bfm.view(ees)).view(ees).matmul(
        afm.unsqueeze(2).unsqueeze(-1)).squeeze()


def ggnn_msg_pass(afm, bfm_int, sd, prefix):
    """GGNNMsgPass.forward (ggnn_msg_pass.py:17-31); bfm_int [B,N,N] int64 bond types, 0 = no bond."""
    B, N, nf = afm.shape
    adj_w = sd[prefix + "adj_w"]
    mf = adj_w.shape[1]
    weights = torch.cat([sd[prefix + "zeros"], adj_w])                                    # :19
    emb = torch.index_select(weights, 0, bfm_int.reshape(-1)).view(B, N, N, mf, nf)       # :20-21
    emb = emb.permute(0, 1, 3, 2, 4).contiguous().view(-1, N * mf, N * nf)                # :22
    return emb.bmm(afm.reshape(B, N * nf, 1)).view(B, N, mf) + sd[prefix + "message_bias"]  # :29-30


# ----------------------------------------------------------------------------------------------
# aggregators
# ----------------------------------------------------------------------------------------------
def adj_msg_agg(messages, adj):
    """adjacent_message_agg.py:18."""
    return messages.mul(adj.unsqueeze(-1)).sum(dim=-2)


def wadj_msg_agg(messages, adj):
    """weighted_adjacent_message_agg.py:20."""
    return messages.mul(torch.softmax(adj, dim=-1).unsqueeze(-1)).sum(dim=-2)


def att_msg_agg(messages, adj, sd, prefix):
    """attention_message_agg.py:10-24 with the default Softmax(dim=-1) over the size-1 axis."""
    w = torch.nn.functional.linear(adj.unsqueeze(-1), sd[prefix + "att.0.weight"], sd[prefix + "att.0.bias"])
    return messages.mul(torch.softmax(w, dim=-1)).sum(dim=-2)


# ----------------------------------------------------------------------------------------------
# update
# ----------------------------------------------------------------------------------------------
def gru_update(messages, node_states, mask, sd, prefix):
    """GRUUpdate.forward -> GRUCell.forward (gru_update.py:26-35, 66-68)."""
    d = node_states.shape[-1]
    m = messages.reshape(-1, messages.shape[-1])
    h = node_states.reshape(-1, d)
    mk = mask.reshape(-1).unsqueeze(-1)
    rzn_i = m.matmul(sd[prefix + "gru_cell.weight_ih"]) + sd[prefix + "gru_cell.bias_ih"]
    rzn_h = h.matmul(sd[prefix + "gru_cell.weight_hh"]) + sd[prefix + "gru_cell.bias_hh"]
    ri, zi, ni = torch.split(rzn_i, d, dim=-1)
    rh, zh, nh = torch.split(rzn_h, d, dim=-1)
    r = torch.sigmoid(ri + rh) * mk
    z = torch.sigmoid(zi + zh) * mk
    n = torch.tanh(ni + r.mul(nh)) * mk
    h_prime = (1 - z).mul(n) + z.mul(h)
    return h_prime.mul(mk).view(node_states.shape)


# ----------------------------------------------------------------------------------------------
# masked batch norms
# ----------------------------------------------------------------------------------------------
def mask_batch_norm(x, mask, eps=1e-6):
    """MaskBatchNorm.forward (mask_batch_norm.py:9-15): unmasked row sum for the mean, eps inside sqrt."""
    mk = mask.reshape(-1).unsqueeze(-1)
    t = x.reshape(-1, x.shape[-1])
    mean = t.sum(dim=0) / mk.sum()
    var = ((t - mean) * mk).pow(2).sum(dim=0) / mk.sum()
    return (((t - mean) * mk) / (var + eps).sqrt()).view(x.shape)


def mask_batch_norm_1d(x, mask, sd, prefix, training=True, momentum=0.1, eps=1e-5, buffers=None):
    """MaskBatchNorm1d.forward (mask_batch_norm.py:20-38): eps OUTSIDE the sqrt, biased running var.

    `buffers` (dict, optional) receives the updated running_mean / running_var under the same keys.
    """
    mk = mask.reshape(-1).unsqueeze(-1)
    y = x.reshape(-1, x.shape[-1])
    mean = (y * mk).sum(dim=0) / mk.sum()
    var = ((y - mean) * mk).pow(2).sum(dim=0) / mk.sum()
    rm_key, rv_key = prefix + "running_mean", prefix + "running_var"
    src = buffers if buffers is not None and rm_key in buffers else sd
    if not training:
        y = (y - src[rm_key]) / (src[rv_key] ** .5 + eps)
    else:
        if buffers is not None:
            with torch.no_grad():
                buffers[rm_key] = (1 - momentum) * src[rm_key] + momentum * mean
                buffers[rv_key] = (1 - momentum) * src[rv_key] + momentum * var
        y = (y - mean) / (var.sqrt() + eps)
    y = sd[prefix + "weight"] * y + sd[prefix + "bias"]
    return (y * mk).view(x.shape)


# ----------------------------------------------------------------------------------------------
# readouts
# ----------------------------------------------------------------------------------------------
def graph_level_output(x, mask, sd, prefix):
    """GraphLevelOutput.forward (graph_level_output.py:30-47); x is cat([node_state, afm])."""
    lin = torch.nn.functional.linear
    wi, bi = sd[prefix + "i.0.weight"], sd[prefix + "i.0.bias"]
    wj, bj = sd[prefix + "j.0.weight"], sd[prefix + "j.0.bias"]
    if mask is not None:
        xm = x * mask
        g = torch.softmax(lin(xm, wi, bi), dim=-1) * lin(xm, wj, bj) * mask               # :36
    else:
        g = torch.softmax(lin(x, wi, bi).sum(dim=1), dim=-1).unsqueeze(1) * lin(x, wj, bj)  # :39
    return g.sum(dim=1)                                                                    # :47


def graph_level_output_atoms(x, mask, sd, prefix):
    """GraphLevelOutput with the commented `return gated_activations` (graph_level_output.py:46) instead of :47: the
    per-atom readout [B,N,O] that normed_encoded_basic_model_ecfp.py:70-71 and its driver
    (test_graph_encode_norm_ecfp.py:137) were written against."""
    lin = torch.nn.functional.linear
    xm = x * mask
    return (torch.softmax(lin(xm, sd[prefix + "i.0.weight"], sd[prefix + "i.0.bias"]), dim=-1)
            * lin(xm, sd[prefix + "j.0.weight"], sd[prefix + "j.0.bias"]) * mask)           # :36, :46


def lstm_cell_hidden(hprev, cprev, sd, prefix):
    """LSTMCellHidden.forward (set2vec.py:68-75)."""
    i = torch.sigmoid(hprev.matmul(sd[prefix + "w_hi"]) + sd[prefix + "b_hi"])
    f = torch.sigmoid(hprev.matmul(sd[prefix + "w_hf"]) + sd[prefix + "b_hf"])
    g = torch.tanh(hprev.matmul(sd[prefix + "w_hg"]) + sd[prefix + "b_hg"])
    o = torch.sigmoid(hprev.matmul(sd[prefix + "w_ho"]) + sd[prefix + "b_ho"])
    cprime = f * cprev + i * g
    return o * torch.tanh(cprime), cprime


def set2vec(x, mask, sd, prefix, steps=100, mprev=None, cprev=None):
    """Set2Vec.forward, inner_prod="default" (set2vec.py:93-151); softmax over ALL B*N rows (:139)."""
    B, N, F = x.shape
    if mprev is None:
        mprev = torch.zeros(B, F, dtype=x.dtype)                                          # :110-111
    mprev = torch.cat([mprev, torch.zeros(B, F, dtype=x.dtype)], dim=1)                   # :113
    if cprev is None:
        cprev = torch.zeros(B, F, dtype=x.dtype)                                          # :114-116
    neg = (1 - mask) * _BIG_NEGATIVE if mask is not None else None                        # :120-121
    m = mprev
    for _ in range(steps):
        m, c = lstm_cell_hidden(mprev, cprev, sd, prefix + "lstmcell.")                   # :126
        query = torch.nn.functional.linear(m, sd[prefix + "q_attn.weight"]).unsqueeze(1)  # :128
        energies = torch.nn.functional.linear(torch.tanh(query + x).view(-1, F),
                                              sd[prefix + "e_attn.weight"])               # :131
        if neg is not None:
            energies = energies + neg.reshape(-1, 1)                                      # :137
        att = torch.softmax(energies, dim=0).view(B, -1, 1)                               # :139
        read = att.mul(x).sum(dim=1)                                                      # :142
        m = torch.cat([m, read], dim=1)                                                   # :144
        mprev, cprev = m, c
    return m


# ----------------------------------------------------------------------------------------------
# model loops (callers; reference models/*.py) -- functional, state_dict keyed like the reference
# ----------------------------------------------------------------------------------------------
def lipo_model(afm, bfm, adj, mask, sd, prefix="", steps=6, buffers=None, training=True):
    """lipo_basic_model.BasicModel.forward (:81-86): HEAD message form, aggregator never called."""
    mf = sd[prefix + "mf.message_bias"].shape[0]
    node_state = afm
    emb = None
    for i in range(steps):
        if i == 0:  # reuse_graph_tensors=(i != 0) re-uses the cached edge embedding (:85)
            emb = edge_embed_head(afm, bfm, sd, prefix + "mf.", mf)
        messages = edge_network_head(afm, bfm, sd, prefix + "mf.", mf, emb)
        m = mask_batch_norm_1d(messages, mask, sd, prefix + "ma_bn.", training, buffers=buffers)
        node_state = gru_update(m, node_state, mask, sd, prefix + "uf.")
        node_state = mask_batch_norm_1d(node_state, mask, sd, prefix + "bn.", training, buffers=buffers)
    return graph_level_output(torch.cat([node_state, afm], dim=-1), mask, sd, prefix + "of.")


def basic_model(afm, bfm, adj, mask, sd, prefix="", steps=3, chain_state=True):
    """basic_model.BasicModel.forward (:50-58) / basic_graph_autoencoder.Encoder.encode (:34-42,
    chain_state=False: the update always starts from afm) on the documented per-pair messages."""
    mf = sd[prefix + "mf.message_bias"].shape[0]
    node_state = afm
    emb = edge_embed_pairs(afm, bfm, sd, prefix + "mf.", mf)   # cached after step 0 (:57, reuse_graph_tensors)
    for _ in range(steps):
        agg = adj_msg_agg(edge_network_pairs(afm, bfm, sd, prefix + "mf.", mf, emb), adj)
        node_state = gru_update(agg, node_state if chain_state else afm, mask, sd, prefix + "uf.")
    return graph_level_output(torch.cat([node_state, afm], dim=-1), mask, sd, prefix + "of.")


def normed_basic_model(afm, bfm, adj, mask, sd, prefix="", steps=3):
    """normed_basic_model.BasicModel.forward (:56-59): one EdgeNetwork per step + MaskBatchNorm."""
    node_state = afm
    for t in range(steps):
        p = prefix + "mf%d." % t
        mf = sd[p + "message_bias"].shape[0]
        agg = adj_msg_agg(edge_network_pairs(afm, bfm, sd, p, mf), adj)
        node_state = mask_batch_norm(gru_update(agg, node_state, mask, sd, prefix + "uf."), mask)
    return graph_level_output(torch.cat([node_state, afm], dim=-1), mask, sd, prefix + "of.")


def att_model(afm, bfm, adj, mask, sd, prefix="", steps=3, s2v_steps=100, agg="adj"):
    """att_model.BasicModel.forward (:56-59): AttEdgeNetwork, AdjMsgAgg|AttMsgAgg, GRU, MaskBatchNorm, Set2Vec."""
    node_state = afm
    for t in range(steps):
        p = prefix + "mf%d." % t
        mf = sd[p + "message_bias"].shape[0]
        msgs = att_edge_network_pairs(afm, bfm, sd, p, mf)
        a = adj_msg_agg(msgs, adj) if agg == "adj" else att_msg_agg(msgs, adj, sd, prefix + "ma.")
        node_state = mask_batch_norm(gru_update(a, node_state, mask, sd, prefix + "uf."), mask)
    return set2vec(torch.cat([node_state, afm], dim=-1), mask, sd, prefix + "of.", s2v_steps)


def _encoder(x, sd, prefix):
    """AtomAutoEncoder/BondAutoEncoder .encoder (encoders/*_autoencoder.py:7-11): Linear(no bias), Tanh, Linear."""
    h = torch.tanh(torch.nn.functional.linear(x, sd[prefix + "0.weight"]))
    return torch.nn.functional.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


def normed_encoded_ecfp_model(afm, bfm, adj, mask, sd, prefix="", steps=3, buffers=None, training=True):
    """normed_encoded_basic_model_ecfp.BasicModel.forward (:65-71) with the per-atom readout its `obn(output, mask)`
    call needs (graph_level_output.py:46): BASELINE config 4."""
    y = normed_encoded_model(afm, bfm, adj, mask, sd, prefix, steps, buffers, training, ma_bn=False,
                             readout=graph_level_output_atoms)
    return mask_batch_norm_1d(y, mask, sd, prefix + "obn.", training, buffers=buffers)    # :71


def normed_encoded_model(afm, bfm, adj, mask, sd, prefix="", steps=3, buffers=None, training=True, ma_bn=True,
                         readout=None):
    """normed_encoded_basic_model.BasicModel.forward (:67-72; ma_bn=True) and the message-passing part of
    normed_encoded_basic_model_ecfp.BasicModel.forward (:65-70; ma_bn=False -- its trailing `obn` call on the
    [B,O] output, :71, is shape-inconsistent at HEAD, SURVEY §2.3, and is left to the caller)."""
    afm = mask_batch_norm_1d(_encoder(afm, sd, prefix + "ae."), mask, sd, prefix + "aebn.", training, buffers=buffers)
    bfm = mask_batch_norm_1d(_encoder(bfm, sd, prefix + "be."), adj, sd, prefix + "bebn.", training, buffers=buffers)
    node_state = afm
    for t in range(steps):
        p = prefix + "mf%d." % t
        mf = sd[p + "message_bias"].shape[0]
        agg = adj_msg_agg(edge_network_pairs(afm, bfm, sd, p, mf), adj)
        if ma_bn:
            agg = mask_batch_norm_1d(agg, mask, sd, prefix + "ma_bn%d." % t, training, buffers=buffers)
        node_state = mask_batch_norm_1d(gru_update(agg, node_state, mask, sd, prefix + "uf."), mask, sd,
                                        prefix + "bn%d." % t, training, buffers=buffers)
    return (readout or graph_level_output)(torch.cat([node_state, afm], dim=-1), mask, sd, prefix + "of.")


# ----------------------------------------------------------------------------------------------
# edge compaction (integer, bit-exact contract)
# ----------------------------------------------------------------------------------------------
def compact_edges(bfm, adj):
    """The edge set the kernels must reproduce bit-exactly: pairs (b,i,j) with a non-zero
    bond row or a non-zero adjacency entry, in row-major (torch.nonzero) order."""
    B, N = adj.shape[:2]
    keep = (bfm != 0).any(-1).logical_or(adj != 0)
    idx = keep.nonzero()
    dst = idx[:, 0] * N + idx[:, 1]
    src = idx[:, 0] * N + idx[:, 2]
    counts = keep.reshape(B * N, N).sum(-1)
    row_ptr = torch.zeros(B * N + 1, dtype=torch.int64)
    row_ptr[1:] = counts.cumsum(0)
    return dict(row_ptr=row_ptr.to(torch.int32), dst=dst.to(torch.int32), src=src.to(torch.int32),
                w=adj[keep], x=bfm[keep])


def xavier_gain_sigmoid():
    return 1.0  # torch.nn.init.calculate_gain('sigmoid'), gru_update.py:18


def count_params(sd):
    seen, total = set(), 0
    for v in sd.values():
        if v.data_ptr() not in seen:
            seen.add(v.data_ptr())
            total += v.numel()
    return total
