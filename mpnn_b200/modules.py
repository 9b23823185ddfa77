"""Drop-in nn.Modules for the reference's message-passing plug-in API.

Same class names, constructor signatures, forward signatures, parameter names/shapes and state_dict keys
as hochshi/mpnn's `mpnn_functions` package and `models/mask_batch_norm.py`, so the reference's model files
(constructor injection, models/basic_model.py:7-32, or `from mpnn_functions import *`, :3) run against
them unchanged.  Every forward/backward is a call into the sm_100a CUDA library through the C-ABI
(`mpnn_b200._lib`); there is no PyTorch or CPU fallback -- CPU tensors raise.
"""
import torch
from torch import nn

from . import graph
from .functional import (BilinearEdgeFn, ChainFn, DenseAggFn, MultiEdgeNetTableFn, EdgeMessageFn, EdgeNetTableFn, EdgeTrunkFn, GatherRowsFn, GatherSumFn,
                         GraphLevelOutputFn, GRUFn, GRUParamHubFn, SharedGradSession, TableHolder, LinearFn, LSTMCellHiddenFn, MaskBN1dFn, MaskBNFn,
                         Set2VecFn, SoftmaxMulFn,
                         TableLayoutFn, TypedMessageFn, TypedMessageTCFn, chain_supported, table_dp, tc_dp, typed_dp)
from . import _lib
from .functional import _note_forward_side_work, _side_stream, real_rows, TypeGatherFn, TypedRowBNFn
import os
import weakref

SIBLING_PREFETCH = os.environ.get("MPNN_B200_SIBLING_PREFETCH", "1") != "0"
SHARED_GRAD_HUB = os.environ.get("MPNN_B200_SHARED_GRAD_HUB", "1") != "0"
LAZY_CHAIN = os.environ.get("MPNN_B200_LAZY_CHAIN", "1") != "0"   # module calls build a lazy chain (fused T-step kernel)

_N_TIED = 50  # edge_network.py:20


def _key(t):
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape))


# =================================================================================================
# message functions
# =================================================================================================
class _LazyTensor(object):
    """Tensor-like handle whose value is computed on first use (`materialize()`): attribute access, operators and torch
    functions (the `__torch_function__` protocol) all evaluate it first, so arbitrary code can consume it."""

    _value = None

    def materialize(self):
        raise NotImplementedError

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def unwrap(a):
            if isinstance(a, _LazyTensor):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(x) for x in a)
            return a
        kwargs = {k: unwrap(v) for k, v in (kwargs or {}).items()}
        return func(*[unwrap(a) for a in args], **kwargs)

    __hash__ = object.__hash__


def _binop(name):
    def f(self, other):
        if isinstance(other, _LazyTensor):
            other = other.materialize()
        return getattr(self.materialize(), name)(other)
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__matmul__", "__getitem__", "__pow__", "__eq__", "__ne__", "__lt__", "__gt__", "__le__", "__ge__"):
    setattr(_LazyTensor, _n, _binop(_n))
_LazyTensor.__neg__ = lambda self: -self.materialize()
_LazyTensor.__hash__ = object.__hash__


class LazyMessages(_LazyTensor):
    """What EdgeNetwork.forward returns: the messages of one (afm, bfm) pair, not yet evaluated.

    The reference HEAD is internally inconsistent: EdgeNetwork.forward returns already-summed
    messages [B,N,mf] (edge_network.py:50-51) while the aggregators expect per-pair messages [B,N,N,mf]
    (adjacent_message_agg.py:13-18; the commented lines edge_network.py:40,52).  One object serves both:
      * an mpnn_b200 aggregator recognises it and calls `.aggregate(...)`: the documented per-pair form,
        fused with the aggregation, never materialising [B,N,N,mf];
      * any other consumer (e.g. lipo_basic_model.py:85 feeds it straight to a batch norm) gets the HEAD
        tensor [B,N,mf]: attribute access, operators and torch functions all materialise it first.
    """

    def __init__(self, net, afm, bfm, reuse):
        self._net, self._afm, self._bfm, self._reuse = net, afm, bfm, reuse
        self._value = None

    # ---- HEAD form ----
    def materialize(self):
        if self._value is None:
            self._value = self._net._head_messages(self._afm, self._bfm, self._reuse)
        return self._value

    # ---- documented per-pair form, fused with the aggregation ----
    def aggregate(self, adj, alpha_fn=None, gamma_fn=None):
        return self._net._aggregated_messages(self._afm, self._bfm, adj, self._reuse, alpha_fn, gamma_fn)

    def __len__(self):
        return self._afm.shape[0]

    def __repr__(self):
        return "LazyMessages(%s, afm=%s)" % (type(self._net).__name__, tuple(self._afm.shape))


class EdgeNetwork(nn.Module):
    """Gilmer edge network (reference mpnn_functions/message/edge_network.py).

    Parameters are held by real nn.Linear children laid out exactly like the reference's `edge_map`
    (growth layers, the SAME Sequential(Linear(P,P,bias=False), act) object repeated 50 times, last Linear)
    so state_dict keys, `model.apply(init_weights)` and optimisers behave identically; the kernels read the
    weights in place.
    """

    def __init__(self, node_features, edge_features, message_features, activation_fn=None, attn_act=None):
        super(EdgeNetwork, self).__init__()
        self.nf = node_features
        self.ef = edge_features
        self.mf = message_features
        self.act_fn = activation_fn if activation_fn is not None else nn.ReLU()
        if not isinstance(self.act_fn, nn.ReLU):
            raise NotImplementedError("mpnn_b200.EdgeNetwork: only ReLU is implemented natively (no fallback); got %r"
                                      % (self.act_fn,))
        edge_map = []
        in_layer = self.ef
        self._growth_idx = []
        while in_layer ** 2 < self.nf * self.mf:
            self._growth_idx.append(len(edge_map))
            edge_map.append(nn.Linear(in_layer, in_layer ** 2))
            edge_map.append(self.act_fn)
            in_layer = in_layer ** 2
        self.P = in_layer
        self._tied_idx = len(edge_map)
        edge_map += [nn.Sequential(nn.Linear(in_layer, in_layer, bias=False), self.act_fn)] * _N_TIED
        self._last_idx = len(edge_map)
        edge_map.append(nn.Linear(in_layer, self.nf * self.mf))
        self.edge_map = nn.Sequential(*edge_map)
        self.message_bias = nn.Parameter(torch.zeros(self.mf))
        self._table_prefetch = None   # (edge-list, table, tableT, event): computed early by a sibling (see _table)
        self._table_group = None      # weakrefs of the networks that shared an edge list, in order of first use
        self._trunk_cache = None   # (edge-list, bfm key, X)         <- the reference's self.edge_embed
        self._table_cache = None   # (edge-list, table, tableT)      <- same role on the typed path
        self._msg_cache = {}       # message tensors of the current edge embedding

    # ---- pieces -----------------------------------------------------------------------------------
    def _trunk(self, bfm, el, reuse):
        """x = edge_map[:-1](bond rows) on the compacted rows (+ zero row), cached like self.edge_embed."""
        if el.E is None:
            raise RuntimeError("mpnn_b200.EdgeNetwork: this configuration (differentiable bond features, attention "
                               "gate, or > 32 features) is not available in capacity (graph-capture) mode")
        c = self._trunk_cache
        if reuse and c is not None and c[0] is el:
            return c[2]
        rows = graph.GatherEdgeRows.apply(bfm, el) if bfm.requires_grad else el.rows
        gw = [self.edge_map[i].weight for i in self._growth_idx]
        gb = [self.edge_map[i].bias for i in self._growth_idx]
        X = EdgeTrunkFn.apply(rows, self.edge_map[self._tied_idx][0].weight, _N_TIED, *(gw + gb))
        self._trunk_cache = (el, None, X)
        self._msg_cache = {}
        return X

    def _last(self):
        last = self.edge_map[self._last_idx]
        return last.weight, last.bias

    # ---- typed path: edge network once per DISTINCT bond row -> table of matrices (csrc/typed.cu) ----------
    _typed_capable = True      # subclasses that gate the sender state per pair turn this off

    def _typed_ok(self, bfm, el):
        """The typed path serves the plain edge network on data bond features (the reference's datasets):
        bfm is not differentiated, the batch holds few distinct bond rows, and the feature widths are <= 32
        (CUDA-core gather kernels, csrc/typed.cu) or 33..256 in multiples of 4 (tcgen05 grouped GEMM,
        csrc/tc_message.cu)."""
        if isinstance(bfm, graph.TypedBonds):     # distinct rows are differentiable inputs of the table
            return True
        if not self._typed_capable or bfm.requires_grad or table_dp(self.nf, self.mf) < 0:
            return False
        ti = el.typed()
        return ti.type_ptr is not None

    def _table(self, el, reuse):
        c = self._table_cache
        if reuse and c is not None and c[0] is el:
            return c[1], c[2]
        pf = self._table_prefetch
        self._table_prefetch = None
        if pf is not None and pf[0] is el:
            # computed ahead of time on the side stream when a sibling network first saw this edge list
            torch.cuda.current_stream().wait_event(pf[3])
            table, tableT = pf[1], pf[2]
            table.record_stream(torch.cuda.current_stream())
            tableT.record_stream(torch.cuda.current_stream())
        else:
            table, tableT = self._compute_table(el)
        self._table_cache = (el, table, tableT)
        self._msg_cache = {}
        self._note_table_user(el)
        return table, tableT

    # ---- sibling prefetch ---------------------------------------------------------------------------------------
    # Models with one EdgeNetwork per step (normed_basic_model.py:24-27, att_model.py) call mf_0, mf_1, ... on the SAME
    # bond tensor; each table depends only on the distinct bond rows and that network's weights, and each one is a
    # latency-bound chain of 52 layers.  The order of first uses of an edge list is remembered as a group; on later
    # batches the first member to arrive enqueues the tables of the others on the side stream (a parallel branch
    # under CUDA-graph capture), so steps 2..T find their table ready.  A table that ends up unused is just dropped.
    def _note_table_user(self, el):
        users = el.__dict__.setdefault("_table_users", [])
        first = not users
        users.append(weakref.ref(self))
        if len(users) >= 2:
            for r in users:
                n = r()
                if n is not None:
                    n._table_group = users
        elif first and SIBLING_PREFETCH and self._table_group is not None:
            peers = [r() for r in self._table_group]
            peers = [n for n in peers if n is not None and n is not self and n._typed_capable]
            if peers and self._table_group[0]() is self:
                self._prefetch_tables(el, peers)

    def _prefetch_tables(self, el, peers):
        main = torch.cuda.current_stream()
        _, side = _side_stream(el.row_ptr.device)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            for n in peers:
                table, tableT = n._compute_table(el)
                done = torch.cuda.Event()
                done.record(side)
                n._table_prefetch = (el, table, tableT, done)
        el.typed().urows.record_stream(side)
        _note_forward_side_work(el.row_ptr.device)

    def _compute_table(self, el):
        ti = el.typed()
        gw = [self.edge_map[i].weight for i in self._growth_idx]
        gb = [self.edge_map[i].bias for i in self._growth_idx]
        w_tied = self.edge_map[self._tied_idx][0].weight
        W, Bv = self._last()
        lib = _lib.load()
        if table_dp(self.nf, self.mf) <= lib.mpnn_enet_max_dp() and lib.mpnn_enet_supported(self.ef, len(gw), self.P):
            table, tableT = EdgeNetTableFn.apply(ti.urows, w_tied, _N_TIED, W, Bv, self.nf, self.mf, *(gw + gb))
            table._mpnn_holder = TableHolder()
        else:   # wide trunks (P = 256, 625, 4096): generic trunk + last Linear on the distinct rows
            X = EdgeTrunkFn.apply(ti.urows, w_tied, _N_TIED, *(gw + gb))
            flat = LinearFn.apply(X[:, :self.P].contiguous(), W, Bv)
            table, tableT = TableLayoutFn.apply(flat, self.nf, self.mf)
        return table, tableT

    def _head_messages_tc(self, afm, table, el):
        """HEAD form (edge_network.py:50-51) around the tensor-core grouped GEMM: every pair (i, j) of a graph
        contributes A(bfm[b,i,j]) h_j and non-bonded pairs share the zero-row matrix T0, so
        M[i] = sum_{e in E(i)} (T[u_e] - T0) h[src_e]  +  T0 S[b]  +  beta,   S[b] = sum_j h[b,j]."""
        B, N, nf = afm.shape
        z = el.typed().zero_type
        T0 = table[z]                                            # [DP, DP], T0[l][k]
        table_h = table - T0.unsqueeze(0)
        tableT_h = table_h.detach().transpose(1, 2).contiguous()
        M = TypedMessageTCFn.apply(afm.reshape(-1, nf), table_h, tableT_h, el, False, self.nf, self.mf)
        S = afm.sum(dim=1)                                       # [B, nf]
        base = LinearFn.apply(S, T0[:self.nf, :self.mf].t().contiguous(), self.message_bias)   # [B, mf]
        return M.view(B, N, self.mf) + base.unsqueeze(1)

    def _sender_vectors(self, afm, bfm, el):
        """(G, gather): what multiplies the edge matrix -- the sender state itself for the plain edge network."""
        return afm.reshape(-1, self.nf), True

    def _nonedge_vectors(self, afm, el, H):
        """per-row vector that multiplies A(x_0): sum of sender states over the NON-edge pairs of the row."""
        B, N, nf = afm.shape
        S = afm.sum(dim=1, keepdim=True).expand(B, N, nf).reshape(-1, nf)
        return S - GatherSumFn.apply(H, el.row_ptr, el.edge_src, el.n_rows, el.col_ptr, el.csc_dst)

    def _head_messages(self, afm, bfm, reuse):
        """edge_network.py:42-51: out[b,i] = sum over ALL j of A(bfm[b,i,j]) afm[b,j] + message_bias."""
        B, N, nf = afm.shape
        el = graph.edge_list_for(bfm, None)
        k = ("head", _key(afm))
        if self._typed_ok(bfm, el):
            table, tableT = self._table(el, reuse)
            if reuse and k in self._msg_cache:
                return self._msg_cache[k]
            if typed_dp(self.nf, self.mf) >= 0:
                M = TypedMessageFn.apply(afm.reshape(-1, nf), table, tableT, self.message_bias, el, None, True,
                                         self.nf, self.mf, getattr(table, "_mpnn_holder", None)).view(B, N, self.mf)
            else:
                M = self._head_messages_tc(afm, table, el)
            self._msg_cache[k] = M
            return M
        X = self._trunk(bfm, el, reuse)
        if reuse and k in self._msg_cache:
            return self._msg_cache[k]
        if bfm.requires_grad:
            raise NotImplementedError("mpnn_b200.EdgeNetwork: HEAD-form messages with a differentiable bfm are not "
                                      "implemented (the non-bonded rows' gradient); use an aggregator")
        G, gather = self._sender_vectors(afm, bfm, el)
        Q = self._nonedge_vectors(afm, el, afm.reshape(-1, nf))
        W, Bv = self._last()
        M = EdgeMessageFn.apply(X, G, None, Q, W, Bv, self.message_bias, el, gather, self.nf, self.mf, self.P)
        M = M.view(B, N, self.mf)
        self._msg_cache[k] = M
        return M

    def _aggregated_messages(self, afm, bfm, adj, reuse, alpha_fn, gamma_fn):
        """sum_j weight[b,i,j] * (A(bfm[b,i,j]) g[b,i,j])  (edge_network.py:52 + the aggregator), no bias."""
        B, N, nf = afm.shape
        el = graph.edge_list_for(bfm, adj)
        k = ("agg", _key(afm), _key(adj), id(alpha_fn), id(gamma_fn))
        if alpha_fn is None and gamma_fn is None and self._typed_ok(bfm, el):
            table, tableT = self._table(el, reuse)
            if reuse and k in self._msg_cache:
                return self._msg_cache[k]
            if typed_dp(self.nf, self.mf) >= 0:
                G, gather = self._sender_vectors(afm, bfm, el)   # node states, or one gated vector per edge
                M = TypedMessageFn.apply(G, table, tableT, None, el if gather else el.per_edge_view(), el.edge_w,
                                         False, self.nf, self.mf, getattr(table, "_mpnn_holder", None)
                                         ).view(B, N, self.mf)
            else:
                M = TypedMessageTCFn.apply(afm.reshape(-1, nf), table, tableT, el, True, self.nf,
                                           self.mf).view(B, N, self.mf)
            self._msg_cache[k] = M
            return M
        X = self._trunk(bfm, el, reuse)
        if reuse and alpha_fn is None and k in self._msg_cache:
            return self._msg_cache[k]
        G, gather = self._sender_vectors(afm, bfm, el)
        alpha = el.edge_w if alpha_fn is None else alpha_fn(el)
        Q = None
        if gamma_fn is not None:  # aggregators that also weight the non-bonded pairs
            Q = gamma_fn(el) * self._nonedge_vectors(afm, el, afm.reshape(-1, nf))
        W, Bv = self._last()
        M = EdgeMessageFn.apply(X, G, alpha, Q, W, Bv, None, el, gather, self.nf, self.mf, self.P)
        M = M.view(B, N, self.mf)
        if alpha_fn is None:
            self._msg_cache[k] = M
        return M

    def forward(self, afm, bfm, reuse_graph_tensors=False):
        if not afm.is_cuda:
            raise RuntimeError("mpnn_b200.EdgeNetwork: CUDA tensors required (there is no CPU fallback)")
        if isinstance(bfm, DeferredRows):
            bfm = bfm.materialize()
        if isinstance(bfm, graph.TypedBonds) and not (type(self)._typed_capable and type(self)._typed_ok is
                                                      EdgeNetwork._typed_ok and table_dp(self.nf, self.mf) >= 0):
            bfm = bfm.dense()
        return LazyMessages(self, afm, bfm, bool(reuse_graph_tensors))


class AttEdgeNetwork(EdgeNetwork):
    """reference att_edge_network.py: the sender state is gated, per pair, by softmax_features(attn(cat(h_i, bond)))."""

    def _typed_ok(self, bfm, el):
        # the gate brings one sender vector per edge: served by the table kernels of csrc/typed.cu (widths <= 32)
        return typed_dp(self.nf, self.mf) >= 0 and super(AttEdgeNetwork, self)._typed_ok(bfm, el)

    def __init__(self, node_features, edge_features, message_features, activation_fn=None, attn_act=None):
        super(AttEdgeNetwork, self).__init__(node_features, edge_features, message_features, activation_fn)
        self.attn = nn.Linear(self.nf + self.ef, self.nf)
        self.attn_act = attn_act if attn_act is not None else nn.Softmax(dim=-1)
        if not (isinstance(self.attn_act, nn.Softmax) and self.attn_act.dim in (-1, 3)):
            raise NotImplementedError("mpnn_b200.AttEdgeNetwork: only the default Softmax(dim=-1) gate is implemented")

    def _gate_logits(self, Hr, Xe):
        Wa, ba = self.attn.weight, self.attn.bias
        logits = LinearFn.apply(Hr, Wa[:, :self.nf].contiguous(), ba)
        if Xe is not None:
            logits = logits + LinearFn.apply(Xe, Wa[:, self.nf:].contiguous(), None)
        return logits

    def _sender_vectors(self, afm, bfm, el):
        H = afm.reshape(-1, self.nf)
        el.neutralise_tail()     # (capacity mode: the slots behind the last edge get valid indices, weight 0)
        Hr = GatherRowsFn.apply(H, el.edge_dst, el.row_ptr, None)          # receiver state per edge
        Hs = GatherRowsFn.apply(H, el.edge_src, el.col_ptr, el.csc_eid)    # sender state per edge
        if el.E is None:
            # capacity (graph-capture) mode: one vector per edge SLOT (graph.PerEdgeView); the bond part of the gate is
            # evaluated on the distinct bond rows and handed out by type, its gradient summed over the type lists
            ti = el.typed()
            Wa, ba = self.attn.weight, self.attn.bias
            LU = LinearFn.apply(ti.urows, Wa[:, self.nf:].contiguous(), None)
            logits = LinearFn.apply(Hr, Wa[:, :self.nf].contiguous(), ba) + TypeGatherFn.apply(LU, ti)
            return SoftmaxMulFn.apply(logits, Hs), False
        rows = graph.GatherEdgeRows.apply(bfm, el) if bfm.requires_grad else el.rows
        return SoftmaxMulFn.apply(self._gate_logits(Hr, rows[:el.E]), Hs), False

    def _nonedge_vectors(self, afm, el, H):
        base = super(AttEdgeNetwork, self)._nonedge_vectors(afm, el, H)
        return SoftmaxMulFn.apply(self._gate_logits(H, None), base)       # gate of a zero bond row

    def _head_messages(self, afm, bfm, reuse):
        raise NotImplementedError("mpnn_b200.AttEdgeNetwork: per-pair messages must be consumed by an mpnn_b200 "
                                  "aggregator (AdjMsgAgg / WAdjMsgAgg / AttMsgAgg); the reference's own HEAD forward "
                                  "fails for this class (SURVEY.md 2.3)")


class BiLiniearEdgeNetwork(nn.Module):
    """reference bilinear_edge_network.py: parameter-free message h_j^T E_ij h_i with the bond row viewed as an
    [nf, nf, nf] tensor (edge_features == node_features**3; message width == node_features).  Returns the same lazy
    handle as EdgeNetwork: an mpnn_b200 aggregator consumes the per-pair messages on the compacted edge list, any
    other consumer gets the reference's dense [B, N, N, nf] tensor."""

    def __init__(self, node_features, edge_features, message_features, activation_fn=None, attn_act=None):
        super(BiLiniearEdgeNetwork, self).__init__()
        self.nf, self.ef, self.mf = node_features, edge_features, message_features
        self.act_fn = activation_fn if activation_fn is not None else nn.ReLU()   # unused, as in the reference

    def _edge_messages(self, afm, bfm, adj):
        B, N, nf = afm.shape
        if bfm.shape[-1] != nf ** 3:
            raise RuntimeError("BiLiniearEdgeNetwork: edge_features (%d) must equal node_features**3 (%d) "
                               "(reference bilinear_edge_network.py:31-36)" % (bfm.shape[-1], nf ** 3))
        el = graph.edge_list_for(bfm, adj)
        if el.E is None:
            raise RuntimeError("mpnn_b200.BiLiniearEdgeNetwork is not available in capacity (graph-capture) mode")
        rows = graph.GatherEdgeRows.apply(bfm, el) if bfm.requires_grad else el.rows
        return el, BilinearEdgeFn.apply(afm.reshape(-1, nf), rows, el)

    def _aggregated_messages(self, afm, bfm, adj, reuse, alpha_fn, gamma_fn):
        """sum_j weight[b,i,j] * msg[b,i,j]; pairs without a bond row contribute exactly 0 whatever their weight"""
        B, N, nf = afm.shape
        el, Y = self._edge_messages(afm, bfm, adj)
        alpha = el.edge_w if alpha_fn is None else alpha_fn(el)
        # backward of the CSR sum: edge e receives the gradient of its receiver row (one-entry transposed lists)
        one = torch.arange(el.E + 1, dtype=torch.int32, device=afm.device)
        M = GatherSumFn.apply(Y * alpha.unsqueeze(1), el.row_ptr, None, el.n_rows, one, el.edge_dst)
        return M.view(B, N, nf)

    def _head_messages(self, afm, bfm, reuse):
        """the reference's own return value: dense per-pair messages [B, N, N, nf]"""
        B, N, nf = afm.shape
        el, Y = self._edge_messages(afm, bfm, None)
        pair = el.edge_dst.long() * N + (el.edge_src.long() % N)
        dense = torch.zeros(B * N * N, nf, dtype=torch.float32, device=afm.device)
        return dense.index_put((pair,), Y).view(B, N, N, nf)

    def forward(self, afm, bfm, reuse_graph_tensors=False):
        if not afm.is_cuda:
            raise RuntimeError("mpnn_b200.BiLiniearEdgeNetwork: CUDA tensors required (there is no CPU fallback)")
        bfm = _as_dense(bfm)
        return LazyMessages(self, afm, bfm, bool(reuse_graph_tensors))


class GGNNMsgPass(nn.Module):
    """reference ggnn_msg_pass.py: integer bond type -> [mf, nf] matrix lookup; type 0 (no bond) maps to the zero matrix."""

    def __init__(self, node_features, edge_features, message_features):
        super(GGNNMsgPass, self).__init__()
        self.nf, self.ef, self.mf = node_features, edge_features, message_features
        self.register_parameter('adj_w', nn.Parameter(torch.Tensor(self.ef, self.mf, self.nf)))
        self.register_parameter('message_bias', nn.Parameter(torch.zeros(self.mf).float()))
        self.register_parameter('zeros', nn.Parameter(torch.zeros(1, self.mf, self.nf).float(), requires_grad=False))

    def init_weights(self):
        torch.nn.init.kaiming_uniform_(self.adj_w, nonlinearity='relu')

    def forward(self, afm, bfm, reuse_graph_tensors=False):
        """ggnn_msg_pass.py:17-31: out[b,i] = sum_j adj_w[type(b,i,j) - 1] afm[b,j] + message_bias, type 0 = no bond.
        The integer bond type IS the type id of the typed message path: the pairs with a bond are compacted, the
        parameter tensor is the table of per-type matrices (no dense one-hot tensor, no de-duplication pass)."""
        if not afm.is_cuda:
            raise RuntimeError("mpnn_b200.GGNNMsgPass: CUDA tensors required (there is no CPU fallback)")
        B, N, nf = afm.shape
        DP = table_dp(self.nf, self.mf)
        if DP < 0:
            return self._forward_onehot(afm, bfm)
        el = graph.int_typed_edge_list(bfm, self.ef)
        F = torch.nn.functional
        table = F.pad(self.adj_w.permute(0, 2, 1), (0, DP - self.mf, 0, DP - self.nf, 0, 1))     # T[u][l][k], row ef = 0
        tableT = F.pad(self.adj_w.detach(), (0, DP - self.nf, 0, DP - self.mf, 0, 1))
        H = afm.reshape(-1, nf)
        if typed_dp(self.nf, self.mf) >= 0:
            M = TypedMessageFn.apply(H, table, tableT, self.message_bias, el, None, False, self.nf, self.mf, None)
        else:
            M = TypedMessageTCFn.apply(H, table, tableT, el, False, self.nf, self.mf) + self.message_bias
        return M.view(B, N, self.mf)

    def _forward_onehot(self, afm, bfm):
        B, N, nf = afm.shape
        onehot = torch.zeros(B, N, N, self.ef, dtype=torch.float32, device=afm.device)
        onehot.scatter_(-1, (bfm.long().clamp(min=1) - 1).unsqueeze(-1), (bfm > 0).float().unsqueeze(-1))
        el = graph.edge_list_for(onehot, None)
        W = self.adj_w.permute(1, 2, 0).reshape(self.mf * self.nf, self.ef)   # [(k*nf+l), p]
        Bv = torch.zeros(self.mf * self.nf, dtype=torch.float32, device=afm.device)
        M = EdgeMessageFn.apply(el.rows, afm.reshape(-1, nf), None, None, W, Bv, self.message_bias, el, True,
                                self.nf, self.mf, self.ef)
        return M.view(B, N, self.mf)


# =================================================================================================
# encoders on categorical bond tensors (reference mpnn_functions/encoders/*.py)
# =================================================================================================
class DeferredRows(_LazyTensor):
    """What a `RowwiseSequential` returns for a dense 4-D CUDA data tensor: "module(x)", not yet evaluated.

    The unchanged reference models run `bfm = self.bebn(self.be(bfm), adj)` (normed_encoded_basic_model.py:68): the
    encoder sees the bond tensor before the adjacency is known.  The masked batch norm that receives this handle
    resolves it: it de-duplicates the RAW bond rows keyed with the adjacency value (`graph.typed_bonds`), pushes the few
    distinct rows through the encoder and normalises them in row space (`graph.TypedBonds`); nothing of size B N^2 F is
    read after compaction.  Any other consumer gets the dense result of the stock layers."""

    def __init__(self, module, x):
        self._module, self._x = module, x
        self._value = None

    def materialize(self):
        if self._value is None:
            self._value = nn.Sequential.forward(self._module, self._x)
        return self._value

    def resolve(self, adj):
        """-> TypedBonds holding module(distinct rows), or the dense tensor when the batch is not categorical"""
        x = self._x
        if (self._value is None and torch.is_tensor(adj) and adj.is_cuda and adj.dim() == 3
                and tuple(adj.shape) == tuple(x.shape[:3])):
            tb = graph.typed_bonds(x, adj)
            if isinstance(tb, graph.TypedBonds):
                return nn.Sequential.forward(self._module, tb)
        return self.materialize()

    @property
    def shape(self):
        return torch.Size(tuple(self._x.shape[:-1]) + (self._module.out_features(self._x.shape[-1]),))


class RowwiseSequential(nn.Sequential):
    """nn.Sequential of row-wise layers (Linear / activations).  On a dense [B,N,N,F] CUDA data tensor it returns a
    `DeferredRows` handle (see there); on anything else it is the stock nn.Sequential."""

    def out_features(self, f):
        for m in self:
            if isinstance(m, nn.Linear):
                f = m.out_features
        return f

    def forward(self, x):
        if (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and not x.requires_grad and x.dtype == torch.float32
                and x.shape[1] == x.shape[2]):
            return DeferredRows(self, x)
        return super(RowwiseSequential, self).forward(x)


class Autoencoder(nn.Module):
    """reference mpnn_functions/encoders/auto_encoder.py (same layers and state_dict keys)"""

    def __init__(self, in_dim=784, mid_dim=400, e_dim=20):
        super(Autoencoder, self).__init__()
        self.encoder = RowwiseSequential(nn.Linear(in_dim, mid_dim, bias=False), nn.Sigmoid(),
                                         nn.Linear(mid_dim, e_dim, bias=False), nn.Sigmoid())
        self.decoder = nn.Sequential(nn.Linear(e_dim, mid_dim, bias=False), nn.Sigmoid(),
                                     nn.Linear(mid_dim, in_dim, bias=False), nn.Sigmoid())

    def forward(self, x):
        return self.decoder(self.encoder(x))


class _TanhAutoEncoder(nn.Module):
    def __init__(self, in_f, mid_f, out_f):
        super(_TanhAutoEncoder, self).__init__()
        self.encoder = RowwiseSequential(nn.Linear(in_f, mid_f, bias=False), nn.Tanh(), nn.Linear(mid_f, out_f))
        self.decoder = nn.Sequential(nn.BatchNorm1d(out_f), nn.Linear(out_f, mid_f), nn.Tanh(),
                                     nn.Linear(mid_f, in_f), nn.Sigmoid())

    def forward(self, x):
        return self.decoder(self.encoder(x))


class AtomAutoEncoder(_TanhAutoEncoder):
    """reference mpnn_functions/encoders/atom_autoencoder.py: 30 -> 15 -> 8"""

    def __init__(self):
        super(AtomAutoEncoder, self).__init__(30, 15, 8)


class BondAutoEncoder(_TanhAutoEncoder):
    """reference mpnn_functions/encoders/bond_autoencoder.py: 8 -> 4 -> 2"""

    def __init__(self):
        super(BondAutoEncoder, self).__init__(8, 4, 2)


# =================================================================================================
# aggregators
# =================================================================================================
def _as_dense(messages):
    if isinstance(messages, graph.TypedBonds):
        return messages.dense()
    return messages.materialize() if isinstance(messages, _LazyTensor) else messages


def _typed_mask_bn(tb, mask, bn1d, module=None, eps=1e-6):
    """Adjacency-masked batch norm of a TypedBonds tensor in row space: every sum over the B N^2 rows of the dense
    tensor is a count-weighted sum over the distinct rows (count c_u, mask value a_u) -- one launch each way
    (`functional.TypedRowBNFn`, csrc/bn.cu k_row_bn_*).  bn1d: MaskBatchNorm1d `module` (mask_batch_norm.py:20-38), else
    MaskBatchNorm (:9-15) with `eps`.  None when the mask is not the adjacency the rows were typed with."""
    if graph._key(mask) != tb._adj_key:
        return None
    if not bn1d:
        y = TypedRowBNFn.apply(tb._rows, tb._a, tb._cnt, None, None, None, None, False, True, 0.0, eps)
        return tb.with_rows(y)
    m = module
    training = m.training or not m.track_running_stats
    if m.momentum is None and training and m.track_running_stats:
        return None          # cumulative moving average: not served in row space (dense kernels)
    track = m.track_running_stats
    y = TypedRowBNFn.apply(tb._rows, tb._a, tb._cnt, m.weight if m.affine else None, m.bias if m.affine else None,
                           m.running_mean if track else None, m.running_var if track else None, True, training,
                           m.momentum, m.eps)
    return tb.with_rows(y)


class LazyAgg(_LazyTensor):
    """`AdjMsgAgg()(mf(afm, bfm), adj)`, not yet evaluated: consumed by GRUUpdate it becomes part of a fused step
    (csrc/chain.cu); anything else gets the aggregated messages [B, N, mf] (the fused message + aggregation kernel)."""

    def __init__(self, messages, adj):
        self._messages, self._adj = messages, adj

    def materialize(self):
        if self._value is None:
            self._value = self._messages.aggregate(self._adj)
        return self._value


class LazyState(_LazyTensor):
    """Node state after `uf(ma(mf(afm, bfm), adj), prev, mask)` (kind "gru") or after a masked batch norm of such a
    state (kind "bn"), not yet evaluated.  The reference's loops (normed_basic_model.py:56-59, basic_model.py:50-58,
    normed_encoded_basic_model_ecfp.py:67-69) build a chain of these, one link per module call; the first real consumer
    (`torch.cat([node_state, afm])` in front of the readout) evaluates the WHOLE chain as one persistent kernel
    (`functional.ChainFn`).  Chains the fused kernel does not serve are evaluated link by link by the per-module
    kernels, with identical results."""

    def __init__(self, kind, src, prev, mask, module, eps=None):
        self._kind, self._src, self._prev, self._mask, self._module, self._eps = kind, src, prev, mask, module, eps

    # ---- link-by-link evaluation (what the module call would have done) ----
    def _evaluate_link(self):
        if self._kind == "gru":
            prev = _as_dense(self._prev)
            return self._module._run(self._src.materialize(), prev, self._mask)
        return self._module._run(self._src.materialize(), self._mask, *([] if self._eps is None else [self._eps]))

    def materialize(self):
        if self._value is None:
            v = _fused_chain(self)
            self._value = v if v is not None else self._evaluate_link()
        return self._value


def _tables_for_steps(msgs, el):
    """Per-type matrices of every step's edge network for one edge list.  Networks whose cached table cannot be reused
    are evaluated together: K sibling networks with the same layer plan (normed_basic_model.py:24-27) are ONE launch
    each way (`MultiEdgeNetTableFn`); anything else goes through the per-network path."""
    todo, seen = [], set()
    for m in msgs:
        net = m._net
        c = net._table_cache
        if id(net) in seen:
            continue
        if m._reuse and c is not None and c[0] is el:
            continue
        seen.add(id(net))
        todo.append(net)
    lib = _lib.load()
    ti = el.typed()
    plan = lambda n: (n.nf, n.ef, n.mf, n.P, len(n._growth_idx))
    fusable = (len(todo) >= 2 and len(todo) <= lib.mpnn_enet_max_nets() and len({plan(n) for n in todo}) == 1
               and table_dp(todo[0].nf, todo[0].mf) <= lib.mpnn_enet_max_dp()
               and lib.mpnn_enet_supported(todo[0].ef, len(todo[0]._growth_idx), todo[0].P))
    if not fusable:
        for n in todo:
            n._table_prefetch = None
            table, tableT = n._compute_table(el)
            n._table_cache = (el, table, tableT)
            n._msg_cache = {}
        return
    n0 = todo[0]
    G = len(n0._growth_idx)
    params = []
    for n in todo:
        W, Bv = n._last()
        params += [n.edge_map[n._tied_idx][0].weight, W, Bv]
        params += [n.edge_map[i].weight for i in n._growth_idx] + [n.edge_map[i].bias for i in n._growth_idx]
    outs = MultiEdgeNetTableFn.apply(ti.urows, _N_TIED, n0.nf, n0.mf, len(todo), G, *params)
    for k, n in enumerate(todo):
        n._table_prefetch = None
        n._table_cache = (el, outs[2 * k], outs[2 * k + 1])
        n._msg_cache = {}


AGG_IN_GRU = os.environ.get("MPNN_B200_AGG_IN_GRU", "1") != "0"
REAL_ROWS_LATE = os.environ.get("MPNN_B200_REAL_ROWS_LATE", "1") != "0"
WIDE_COMPACT = os.environ.get("MPNN_B200_WIDE_COMPACT", "1") != "0"


class _CompactEdges(object):
    """The edge list of a batch re-indexed over its REAL rows (mask != 0): what the tensor-core kernels see in
    `_wide_chain`.  Padded rows have no edges, so the CSR / CSC pointers of the real rows are a gather of the padded
    ones; sender / receiver ids go through the inverse of the row list; edge ids (CSC lists, type grouping) are unchanged."""

    def __init__(self, el, real, cap):
        rows = el.n_rows
        dev = real.device
        self.B, self.N, self.ef, self.E, self.Ecap = el.B, el.N, el.ef, el.E, el.Ecap
        self.n_rows = cap
        self.edge_w, self.csc_eid, self.rows = el.edge_w, el.csc_eid, None
        rl = real.long()
        inv = torch.zeros(rows + 1, dtype=torch.int32, device=dev)
        inv[rl] = torch.arange(cap, dtype=torch.int32, device=dev)
        self.row_ptr = torch.cat([el.row_ptr.index_select(0, rl), el.row_ptr[-1:]]).contiguous()
        self.col_ptr = torch.cat([el.col_ptr.index_select(0, rl), el.col_ptr[-1:]]).contiguous()
        # (capacity mode: slots behind the edge count hold arbitrary bits -- clamp before they are used as indices)
        self.edge_src = inv[el.edge_src.clamp(0, rows).long()].contiguous()
        self.edge_dst = inv[el.edge_dst.clamp(0, rows).long()].contiguous()
        import copy
        self._typed = copy.copy(el.typed())
        self._typed._tc_plan = None        # the plan bakes sender / receiver rows in: rebuilt over the compact ids

    def typed(self):
        return self._typed


_NODE_FLAGS = {}     # device -> persistent {0, 0, overflow, 0} of the compact node lists built in capacity mode


def compact_nodes(afm, mask, el):
    """(compact edge list, real-row list [cap], afm on the real rows [cap, d], mask on them [cap]).  Eager: cap = number
    of real rows (one small device->host read, like the edge count).  Capacity mode (graph capture): cap comes from the
    eager warm-up with 10 % headroom, nothing is read back, an overflow sets the sticky flag `GraphedStep.check` reads."""
    B, N, d = afm.shape
    rows = B * N
    dev = afm.device
    m1 = mask.reshape(-1)
    if graph._CAPACITY is None:
        n_real = int((m1 != 0).sum().item())
        graph.STATS["n_real"] = max(graph.STATS.get("n_real", 0), n_real)
        cap = max(n_real, 1)
        if dev not in _NODE_FLAGS:
            _NODE_FLAGS[dev] = torch.zeros(4, dtype=torch.int32, device=dev)
    else:
        seen = graph.STATS.get("n_real", 0)
        cap = min(rows, int(seen * 1.1) + 64) if seen else rows
        flags = _NODE_FLAGS.get(dev)
        if flags is not None and cap < rows:
            flags[2:3].bitwise_or_(((m1 != 0).sum() > cap).to(torch.int32).reshape(1))
            graph._CAPTURED_COUNTS.append(flags)
        elif cap < rows:
            cap = rows      # no persistent flag to report an overflow with: do not truncate
    real = torch.nonzero_static(m1 != 0, size=cap, fill_value=rows).reshape(-1)     # fill: the appended zero row
    elc = _CompactEdges(el, real, cap)
    afm_c = torch.cat([afm.reshape(-1, d), afm.new_zeros(1, d)]).index_select(0, real)
    mask_c = torch.cat([m1, m1.new_zeros(1)]).index_select(0, real)
    return elc, real, afm_c, mask_c


def _wide_chain(steps, base, afm, mask, el, uf, d):
    """The step chain at tensor-core widths (33..256) on the REAL rows only.  The reference pads every graph to the
    largest one of the batch (data_loader.py:52-59): 40 % of the rows of a ZINC-shaped batch are padding, and every
    node-tensor pass of the message / GRU / batch-norm kernels pays for them.  Here the node tensors are gathered to
    [n_real, d] once, all steps run on the compact tensors (the per-module kernels, unchanged), and the result is
    scattered back into the padded layout once.  Exact: padded rows carry no edges and their states are zero."""
    B, N, _ = afm.shape
    rows = B * N
    dev = afm.device
    elc, real, afm_c, mask_c = compact_nodes(afm, mask, el)
    h = afm_c if base is afm else torch.cat([base.reshape(-1, d), afm.new_zeros(1, d)]).index_select(0, real)
    for gru, bn in steps:
        msg = gru._src._messages
        table, tableT = msg._net._table(el, True)
        # (the CSR sum of the per-edge messages is left to the GRU kernel's operand producer: M is filled there)
        M = TypedMessageTCFn.apply(afm_c, table, tableT, elc, True, d, d, AGG_IN_GRU)
        h = uf.gru_cell(M, h, mask_c)
        if bn is not None:
            if isinstance(bn._module, MaskBatchNorm1d):
                h = bn._module._run(h, mask_c)
            else:
                h = bn._module._run(h, mask_c, 1e-6 if bn._eps is None else bn._eps)
    out = torch.zeros(rows + 1, d, dtype=torch.float32, device=dev).index_copy(0, real, h)
    return out[:rows].view(B, N, d)


def _fused_chain(head):
    """Evaluates the chain ending in `head` with the persistent step kernel when every link fits it; None otherwise."""
    steps, node = [], head
    while isinstance(node, LazyState) and node._value is None:
        bn = None
        if node._kind == "bn":
            bn, node = node, node._src
            if not (isinstance(node, LazyState) and node._kind == "gru" and node._value is None):
                return None
        if not isinstance(node._src, LazyAgg) or node._src._value is not None:
            return None
        steps.append((node, bn))
        node = node._prev
    if not steps:
        return None
    steps.reverse()
    base = _as_dense(node)
    g0 = steps[0][0]
    msg0 = g0._src._messages
    afm, bfm, adj, mask, uf = msg0._afm, msg0._bfm, g0._src._adj, g0._mask, g0._module
    if not (torch.is_tensor(afm) and afm.is_cuda and torch.is_tensor(base) and base.is_cuda and afm.dim() == 3
            and base.shape == afm.shape and afm.dtype == torch.float32 and base.dtype == torch.float32):
        return None
    d = afm.shape[-1]
    wide = False
    if uf.nf != d or uf.mf != d:
        return None
    if not chain_supported(d, len(steps)):
        wide = WIDE_COMPACT and d > 32 and tc_dp(d, d) >= 0     # tensor-core widths: real rows only (see _wide_chain)
        if not wide:
            return None
    for gru, bn in steps:
        m = gru._src._messages
        net = m._net
        if (gru._module is not uf or gru._mask is not mask or m._afm is not afm or m._bfm is not bfm
                or gru._src._adj is not adj or type(net)._aggregated_messages is not EdgeNetwork._aggregated_messages
                or not net._typed_capable or net.nf != d or net.mf != d):
            return None
        if bn is not None and (bn._mask is not mask or (isinstance(bn._module, MaskBatchNorm1d)
                                                        and bn._module.momentum is None)):
            return None
    fork = None
    if not wide and REAL_ROWS_LATE and mask.is_cuda:
        # the list of real rows (tiny kernel, side lane) depends on the mask only: forked from HERE, but enqueued behind the
        # compaction / edge networks, so that in a captured step the critical chain (edge networks -> step kernel) is the
        # first branch created behind the graph's root and the side lane the second
        fork = torch.cuda.Event()
        fork.record(torch.cuda.current_stream(mask.device))
    elif not wide:
        real_rows(mask, side=True)     # tiny kernel on a side lane, overlapped with the compaction / edge networks below
    el = graph.edge_list_for(bfm, adj)
    if not all(gru._src._messages._net._typed_ok(bfm, el) for gru, _ in steps):
        return None
    tables, tablesT, bnspec, affine = [], [], [], []
    _tables_for_steps([gru._src._messages for gru, _ in steps], el)
    if fork is not None:
        real_rows(mask, side=True, fork=fork)
    if wide:
        return _wide_chain(steps, base, afm, mask, el, uf, d)
    for gru, bn in steps:
        m = gru._src._messages
        table, tableT = m._net._table(el, True)      # cached by _tables_for_steps
        tables.append(table)
        tablesT.append(tableT)
        if bn is None:
            bnspec.append(dict(kind=0, training=0, eps=0.0, momentum=0.0, affine=False))
        elif isinstance(bn._module, MaskBatchNorm1d):
            mod = bn._module
            tracked = mod.track_running_stats
            bnspec.append(dict(kind=2, training=mod.training or not tracked, eps=mod.eps, momentum=mod.momentum,
                               affine=mod.affine, running_mean=mod.running_mean if tracked else None,
                               running_var=mod.running_var if tracked else None))
            if mod.affine:
                affine += [mod.weight, mod.bias]
        else:
            bnspec.append(dict(kind=1, training=1, eps=1e-6 if bn._eps is None else bn._eps, momentum=0.0, affine=False))
    W_ih, W_hh, b_ih, b_hh = uf.gru_cell.weight_ih, uf.gru_cell.weight_hh, uf.gru_cell.bias_ih, uf.gru_cell.bias_hh
    B, N, _ = afm.shape
    out = ChainFn.apply(afm.reshape(-1, d), base.reshape(-1, d), mask.reshape(-1), el, bnspec, W_ih, W_hh, b_ih, b_hh,
                        *(tables + tablesT + affine))
    return out.view(B, N, d)


class AdjMsgAgg(nn.Module):
    """reference adjacent_message_agg.py: out[b,i] = sum_j adj[b,i,j] * messages[b,i,j]."""

    def __init__(self, adj_dim, attn_act=None):
        super(AdjMsgAgg, self).__init__()

    def forward(self, messages, adj):
        if isinstance(messages, LazyMessages):
            if LAZY_CHAIN and type(messages._net) is EdgeNetwork and torch.is_tensor(adj) and adj.is_cuda:
                return LazyAgg(messages, adj)      # may become part of a fused step (LazyState)
            return messages.aggregate(adj)
        return DenseAggFn.apply(_as_dense(messages), adj)


class WAdjMsgAgg(nn.Module):
    """reference weighted_adjacent_message_agg.py: weights softmax_j(adj[b,i,:]) -- non-neighbours and padded
    atoms get e^0/Z too."""

    def __init__(self, adj_dim, attn_act=None):
        super(WAdjMsgAgg, self).__init__()

    def forward(self, messages, adj):
        if isinstance(messages, LazyMessages):
            def parts(el):
                # softmax over ALL N columns of the row (non-neighbours and padded atoms carry adj = 0): subtract the
                # row maximum max(0, max_e w_e) like torch.softmax does, so large weighted adjacencies cannot overflow
                w = el.edge_w
                dst = el.edge_dst.long()
                with torch.no_grad():
                    mx = torch.zeros(el.n_rows, dtype=torch.float32, device=w.device)
                    if w.numel():
                        mx.scatter_reduce_(0, dst, w.detach(), "amax", include_self=True)
                ew = torch.exp(w - mx[dst])
                e0 = torch.exp(-mx)                       # weight of a zero entry of the row
                deg = (el.row_ptr[1:] - el.row_ptr[:-1]).float()
                Z = GatherSumFn.apply(ew.unsqueeze(1), el.row_ptr, None, el.n_rows, None, None).squeeze(1) \
                    + (el.N - deg) * e0
                return ew, Z, e0
            return messages.aggregate(
                adj,
                alpha_fn=lambda el: (lambda ew, Z, e0: ew / Z[el.edge_dst.long()])(*parts(el)),
                gamma_fn=lambda el: (lambda ew, Z, e0: (e0 / Z).unsqueeze(1))(*parts(el)))
        B, N, _ = adj.shape
        w = SoftmaxMulFn.apply(adj.reshape(B * N, N), None).view(B, N, N)
        return DenseAggFn.apply(messages, w)


class AttMsgAgg(nn.Module):
    """reference attention_message_agg.py: weights act(Linear(adj[..., None])); with the default
    Softmax(dim=-1) over the size-1 axis every weight is exactly 1 (an unmasked sum)."""

    def __init__(self, adj_dim, attn_act=None):
        super(AttMsgAgg, self).__init__()
        self.adj_dim = adj_dim
        self.att = nn.Sequential(
            nn.Linear(adj_dim, 1),
            attn_act if attn_act is not None else nn.Softmax(dim=-1)
        )

    def forward(self, messages, adj):
        if isinstance(messages, LazyMessages):
            dev = adj.device
            return messages.aggregate(
                adj,
                alpha_fn=lambda el: self.att(el.edge_w.unsqueeze(-1)).squeeze(-1),
                gamma_fn=lambda el: self.att(torch.zeros(1, 1, device=dev)).reshape(1, 1))
        return DenseAggFn.apply(messages, self.att(adj.unsqueeze(-1)).squeeze(-1))


# =================================================================================================
# update
# =================================================================================================
class GRUCell(nn.Module):
    """Parameter holder with the reference's names/shapes/init (gru_update.py:5-24); weights are [in, 3d]."""

    def __init__(self, node_features, message_features):
        super(GRUCell, self).__init__()
        self.nf = node_features
        self.mf = message_features
        self.register_parameter('weight_ih', nn.Parameter(torch.Tensor(self.mf, 3 * self.nf)))
        self.register_parameter('weight_hh', nn.Parameter(torch.Tensor(self.nf, 3 * self.nf)))
        self.register_parameter('bias_ih', nn.Parameter(torch.Tensor(3 * self.mf)))
        self.register_parameter('bias_hh', nn.Parameter(torch.Tensor(3 * self.nf)))
        self.init_params()

    def init_params(self):
        torch.nn.init.xavier_uniform_(self.weight_ih, gain=torch.nn.init.calculate_gain('sigmoid'))
        torch.nn.init.xavier_uniform_(self.weight_hh, gain=torch.nn.init.calculate_gain('sigmoid'))
        nn.init.constant_(self.bias_ih, 0.0)
        nn.init.constant_(self.bias_hh, 0.0)

    _session = None

    def _shared_parameters(self):
        """(W_ih, W_hh, b_ih, b_hh, session): while the same parameter values are applied again and again before a
        backward pass (the T message-passing steps), every call goes through ONE hub node, so the steps' weight-gradient
        partials are reduced once instead of T times + 4 (T-1) autograd accumulations (functional.SharedGradSession)."""
        ps = (self.weight_ih, self.weight_hh, self.bias_ih, self.bias_hh)
        if not (SHARED_GRAD_HUB and torch.is_grad_enabled() and all(p.requires_grad for p in ps)):
            return ps + (None,)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        s = self._session
        # (s.filled: a backward pass left partials in the slab without reaching the hub node -- never reuse it)
        if s is None or s.done or s.filled or s.key != key or s.uses >= 64:
            s = SharedGradSession(key)
            s.handles = GRUParamHubFn.apply(s, *ps)
            self._session = s
        return tuple(s.handles) + (s,)

    def forward(self, messages, node_states, mask):
        W_ih, W_hh, b_ih, b_hh, session = self._shared_parameters()
        return GRUFn.apply(messages, node_states, mask.reshape(-1), W_ih, W_hh, b_ih, b_hh, session)


class GRUUpdate(nn.Module):
    """reference gru_update.py:39-68 (including the swapped GRUCell(mf, nf) construction, :53, which makes the
    module usable only with node_features == message_features)."""

    def __init__(self, node_features, message_features):
        super(GRUUpdate, self).__init__()
        self.nf = node_features
        self.mf = message_features
        self.gru_cell = GRUCell(self.mf, self.nf)

    def forward(self, messages, node_states, mask):
        if self.nf != self.mf:
            raise RuntimeError("GRUUpdate requires node_features == message_features (reference gru_update.py:53)")
        if isinstance(messages, LazyAgg) and (torch.is_tensor(node_states) or isinstance(node_states, LazyState)):
            return LazyState("gru", messages, node_states, mask, self)
        return self._run(_as_dense(messages), _as_dense(node_states), mask)

    def _run(self, messages, node_states, mask):
        h = self.gru_cell(messages.reshape(-1, self.mf), node_states.reshape(-1, self.nf), mask)
        return h.view(node_states.shape)


# =================================================================================================
# masked batch norms  (reference models/mask_batch_norm.py)
# =================================================================================================
class MaskBatchNorm(nn.Module):
    def __init__(self):
        super(MaskBatchNorm, self).__init__()

    def forward(self, tensor, mask, eps=1e-6):
        if isinstance(tensor, LazyState) and tensor._kind == "gru" and tensor._value is None and tensor._mask is mask:
            return LazyState("bn", tensor, None, mask, self, eps)
        return self._run(tensor, mask, eps)

    def _run(self, tensor, mask, eps=1e-6):
        if isinstance(tensor, DeferredRows):
            tensor = tensor.resolve(mask)
        if isinstance(tensor, graph.TypedBonds):
            out = _typed_mask_bn(tensor, mask, False, eps=eps)    # mask_batch_norm.py:11-15 on the distinct rows
            if out is not None:
                return out
        tensor = _as_dense(tensor)
        y = MaskBNFn.apply(tensor.reshape(-1, tensor.shape[-1]), mask.reshape(-1), eps)
        return y.view(tensor.shape)


class MaskBatchNorm1d(nn.BatchNorm1d):
    def _typed_forward(self, tb, mask):
        return _typed_mask_bn(tb, mask, True, module=self)    # mask_batch_norm.py:20-38 on the distinct rows

    def forward(self, tensor, mask):
        if isinstance(tensor, LazyState) and tensor._kind == "gru" and tensor._value is None and tensor._mask is mask:
            return LazyState("bn", tensor, None, mask, self)
        return self._run(tensor, mask)

    def _run(self, tensor, mask):
        if isinstance(tensor, DeferredRows):
            tensor = tensor.resolve(mask)
        if isinstance(tensor, graph.TypedBonds):
            out = self._typed_forward(tensor, mask)
            if out is not None:
                return out
        tensor = _as_dense(tensor)
        training = self.training or not self.track_running_stats
        rm = self.running_mean if self.track_running_stats else None
        rv = self.running_var if self.track_running_stats else None
        y = MaskBN1dFn.apply(tensor.reshape(-1, tensor.shape[-1]), mask.reshape(-1),
                             self.weight if self.affine else None, self.bias if self.affine else None,
                             rm, rv, training, self.momentum, self.eps)
        return y.view(tensor.shape)


# =================================================================================================
# readouts
# =================================================================================================
class GraphLevelOutput(nn.Module):
    """reference readout/graph_level_output.py."""

    def __init__(self, node_features, output_dim, time_steps=100, inner_prod="default", activation_fn=None,
                 attn_act=None, dropout=0):
        super(GraphLevelOutput, self).__init__()
        self.in_dim = node_features
        self.out_dim = output_dim
        self.act_fn = activation_fn() if activation_fn is not None else nn.ReLU()
        self.attn_act = attn_act() if attn_act is not None else nn.Softmax(dim=1)
        self.dropout = dropout
        self.i = nn.Sequential(nn.Linear(2 * self.in_dim, self.out_dim))
        self.j = nn.Sequential(nn.Linear(2 * self.in_dim, self.out_dim))

    def forward(self, input_set, mask=None, mprev=None, cprev=None):
        input_set = _as_dense(input_set)
        return GraphLevelOutputFn.apply(input_set, mask, self.i[0].weight, self.i[0].bias, self.j[0].weight,
                                        self.j[0].bias)


class GraphLevelOutputAtoms(GraphLevelOutput):
    """GraphLevelOutput with the reference's commented `return gated_activations` (graph_level_output.py:46) in place
    of the sum over atoms (:47): the per-atom readout [B, N, O] that normed_encoded_basic_model_ecfp.py:70-71
    (`self.obn(output, mask)`) and its driver (test_graph_encode_norm_ecfp.py:137) were written against.
    Same parameters and state_dict keys; pass it as `readout_func=` to the unchanged ecfp model file."""

    def forward(self, input_set, mask=None, mprev=None, cprev=None):
        input_set = _as_dense(input_set)
        if mask is None:
            raise RuntimeError("mpnn_b200.GraphLevelOutputAtoms: the per-atom readout is defined with a mask "
                               "(graph_level_output.py:33-36)")
        if not input_set.is_cuda:
            raise RuntimeError("mpnn_b200.GraphLevelOutputAtoms: CUDA tensors required (there is no CPU fallback)")
        B, N, F2 = input_set.shape
        m = mask.reshape(B * N, 1)
        xm = input_set.reshape(B * N, F2) * m
        u = LinearFn.apply(xm, self.i[0].weight, self.i[0].bias)
        v = LinearFn.apply(xm, self.j[0].weight, self.j[0].bias)
        return (SoftmaxMulFn.apply(u, v) * m).view(B, N, self.out_dim)


class LSTMCellHidden(nn.Module):
    """Parameter holder of the input-less LSTM (reference set2vec.py:13-66): w_h{i,f,g,o} [hd, cd], b_h* [1, cd]."""

    def __init__(self, hidden_dim, cell_dim, bias=True):
        super(LSTMCellHidden, self).__init__()
        self.hd = hidden_dim
        self.cd = cell_dim
        self.bias = bias
        stdv = 1.0 / (self.hd ** 0.5)
        for g in ("i", "f", "g", "o"):
            self.register_parameter("w_h" + g, nn.Parameter(torch.zeros([self.hd, self.cd]).uniform_(-stdv, stdv)))
        for g in ("i", "f", "g", "o"):
            self.register_parameter("b_h" + g, nn.Parameter(torch.zeros([1, self.cd])))

    def cat_params(self):
        W = torch.cat([self.w_hi, self.w_hf, self.w_hg, self.w_ho], dim=1)
        b = torch.cat([self.b_hi, self.b_hf, self.b_hg, self.b_ho], dim=1).reshape(-1)
        return W, b

    def forward(self, hprev, cprev):
        """set2vec.py:68-75: (hprev [B, hd], cprev [B, cd]) -> (h', c').  Inside Set2Vec the cell is evaluated by the
        fused kernels; this is the stand-alone call of the exported class."""
        W, b = self.cat_params()
        return LSTMCellHiddenFn.apply(hprev, cprev, W, b)


class Set2Vec(nn.Module):
    """reference readout/set2vec.py ("default" inner product; the "dot" variant is broken at HEAD)."""

    def __init__(self, node_features, output_dim, time_steps=100, inner_prod="default", activation_fn=None,
                 attn_act=None, dropout=0):
        super(Set2Vec, self).__init__()
        self.nf = 2 * node_features
        self.steps = time_steps
        self.q_attn = nn.Linear(self.nf, self.nf, bias=False)
        if "default" == inner_prod:
            self.ip = True
            self.e_attn = nn.Linear(self.nf, 1, bias=False)
        elif "dot" == inner_prod:
            raise NotImplementedError('mpnn_b200.Set2Vec: inner_prod="dot" fails in the reference itself; not built')
        else:
            raise ValueError("Invalid inner_prod type: {}".format(inner_prod))
        self.add_module('lstmcell', LSTMCellHidden(self.nf * 2, self.nf))

    def forward(self, input_set, mask=None, mprev=None, cprev=None):
        input_set = _as_dense(input_set)
        W, b = self.lstmcell.cat_params()
        m0 = None
        if mprev is not None:   # set2vec.py:113-114: the caller's [B, F] state is padded with a zero read vector
            m0 = torch.cat([mprev, torch.zeros_like(mprev)], dim=1)
        return Set2VecFn.apply(input_set, mask, W, b, self.q_attn.weight, self.e_attn.weight.reshape(-1), self.steps,
                               m0, cprev)
