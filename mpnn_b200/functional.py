"""torch.autograd glue over the C-ABI kernels.  PyTorch is used for device memory, streams and the autograd
tape only; every forward and backward below is a call into libmpnn_b200.so.  No CPU path exists."""
import ctypes
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, f32c, ptr, ptr_array, stream, workspace


# ------------------------------------------------------------------------------------------------
# weight-gradient work off the critical path
# ------------------------------------------------------------------------------------------------
# The backward of the edge network on the distinct bond rows (50 dependent layers, latency-bound) produces only
# PARAMETER gradients: nothing else in the backward pass waits for it.  It is therefore enqueued on a side stream
# (forked from the current stream by an event, also under CUDA-graph capture, where it becomes a parallel branch
# of the graph) and joined once, at the end of the backward pass, by an autograd-engine callback.
_SIDE_STREAMS = {}
_SIDE_PENDING = {}
SIDE_STREAM_ENABLED = os.environ.get("MPNN_B200_SIDE_STREAM", "1") != "0"
# The backward's side lanes only pay inside a captured step (parallel branches of the CUDA graph, explicit event edges);
# with eager launches the Python dispatch is the bottleneck and the lanes buy nothing, so they are off there by default.
# (Round 2: an in-suite failure of the att_model's input gradient first blamed on the eager lanes was the batch-norm
# workspace bug fixed in csrc/bn.cu `carve`; with MPNN_B200_SIDE_STREAM_EAGER=1 the eager att_model step still differs
# from its captured replay in tests/test_gpu_chain.py, so the eager lanes stay off.)
SIDE_STREAM_EAGER = os.environ.get("MPNN_B200_SIDE_STREAM_EAGER", "0") != "0"


def _side_ok():
    return SIDE_STREAM_ENABLED and (SIDE_STREAM_EAGER or torch.cuda.is_current_stream_capturing())


def _side_stream(device, lane=0):
    """lane 0: edge-network work (forward table prefetch, backward); lane 1: the table half of the message backward"""
    d = device.index if device.index is not None else torch.cuda.current_device()
    k = d if lane == 0 else (d, lane)
    if k not in _SIDE_STREAMS:
        _SIDE_STREAMS[k] = torch.cuda.Stream(device=device)
    return k, _SIDE_STREAMS[k]


def _dev_of(k):
    return k[0] if isinstance(k, tuple) else k


# gradient tensors produced on a side stream: data_ptr -> event recorded behind their last kernel.  A consumer that
# also runs on a side stream waits for THAT event instead of forking from the current position of the main stream
# (autograd runs the edge networks' backward nodes last, long after their inputs were ready).
_READY_EVENTS = {}


def _ready_put(t, ev):
    import weakref
    _READY_EVENTS[t.data_ptr()] = (weakref.ref(t), ev)


def _note_produced(t):
    """record an event behind the kernels enqueued so far on the current stream as the producer of `t`"""
    if _side_ok():
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        _ready_put(t, ev)


READY_EVENTS_ENABLED = os.environ.get("MPNN_B200_READY_EVENTS", "1") != "0"


def _ready_get(t):
    """event behind the producer of `t` -- only while the tensor object it was registered for is still alive (so the
    address cannot have been recycled for something else) and `t` starts at the same address (itself or a view)"""
    ent = _READY_EVENTS.pop(t.data_ptr(), None)
    if ent is None or not READY_EVENTS_ENABLED or not _side_ok():
        return None
    src = ent[0]()
    return ent[1] if (src is not None and src.data_ptr() == t.data_ptr()) else None


_FWD_SIDE = set()


def _note_forward_side_work(device, lane=0):
    """forward work was enqueued on a side stream (sibling table prefetch, type sort): `join_side_streams()` waits for it"""
    if not torch.cuda.is_current_stream_capturing():
        return      # eager mode: consumers wait for the work's own event; only a capture must re-join its forked branches
    d = device.index if device.index is not None else torch.cuda.current_device()
    _FWD_SIDE.add(d if lane == 0 else (d, lane))


def join_side_streams():
    """Current stream waits for everything enqueued so far on the side streams.  Needed only by code that must not
    leave the side stream running: the end of a CUDA-graph capture (graphs.GraphedStep calls it)."""
    for k in list(_FWD_SIDE):
        _FWD_SIDE.discard(k)
        if k in _SIDE_STREAMS:
            torch.cuda.current_stream(_dev_of(k)).wait_stream(_SIDE_STREAMS[k])
    _join_side_streams()


def _join_side_streams():
    for k in list(_SIDE_PENDING):
        ev = _SIDE_PENDING.pop(k)
        torch.cuda.current_stream(_dev_of(k)).wait_event(ev)
    _READY_EVENTS.clear()


class _on_side_stream(object):
    """`with _on_side_stream(device, tensors_read): ...` runs the body's launches on the side stream, after
    everything enqueued so far on the current stream; the join is deferred to the end of the backward pass."""

    def __init__(self, device, tensors, lane=0, after=None):
        self.device, self.tensors = device, [t for t in tensors if t is not None]
        self.lane, self.after = lane, after
        self.done = None

    def __enter__(self):
        self.key, self.side = _side_stream(self.device, self.lane)
        if self.after is not None:       # inputs were produced on another side stream: depend on exactly that
            self.side.wait_event(self.after)
        else:
            main = torch.cuda.current_stream(self.device)
            ev = torch.cuda.Event()
            ev.record(main)
            self.side.wait_event(ev)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        first = not _SIDE_PENDING
        ev = torch.cuda.Event()
        ev.record(self.side)
        self.done = ev
        _SIDE_PENDING[self.key] = ev
        self.ctx.__exit__(*exc)
        for t in self.tensors:
            t.record_stream(self.side)
        if first:
            torch.autograd.Variable._execution_engine.queue_callback(_join_side_streams)
        return False


# ------------------------------------------------------------------------------------------------
# launch-bound ops as CUDA graphs
# ------------------------------------------------------------------------------------------------
# Set2Vec is 100 strictly sequential attention steps (set2vec.py:123-148): ~700 forward and ~1200 backward launches of
# microsecond kernels.  When the step as a whole is not being captured (graphs.GraphedStep), the op captures ITS OWN
# launch sequence once per shape into a CUDA graph over static buffers and replays it (copy in, replay, copy out).
OP_GRAPHS_ENABLED = os.environ.get("MPNN_B200_OP_GRAPHS", "1") != "0"
_OP_GRAPHS = {}


def _graphed(key, run, ins, make_bufs, n_out):
    """run(ins, bufs) enqueues the op's kernels reading `ins` and writing `bufs`; the first n_out bufs are returned."""
    if not OP_GRAPHS_ENABLED or torch.cuda.is_current_stream_capturing():
        bufs = make_bufs()
        run(ins, bufs)
        return bufs[:n_out]
    ent = _OP_GRAPHS.get(key)
    if ent is None:                 # first call of this shape: eager (also the warm-up the capture needs)
        _OP_GRAPHS[key] = "warm"
        bufs = make_bufs()
        run(ins, bufs)
        return bufs[:n_out]
    if ent == "warm":
        s_ins = [t.clone() if t is not None else None for t in ins]
        s_bufs = make_bufs()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run(s_ins, s_bufs)
        ent = (g, s_ins, s_bufs)
        _OP_GRAPHS[key] = ent
    g, s_ins, s_bufs = ent
    for sbuf, t in zip(s_ins, ins):
        if sbuf is not None:
            sbuf.copy_(t)
    g.replay()
    return [b.clone() if b is not None else None for b in s_bufs[:n_out]]


class _inline(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


# ------------------------------------------------------------------------------------------------
# Zero-initialised buffers of a captured step.  Accumulators and "rows the kernel skips are zero" outputs each cost a
# fill launch per step; inside a captured step (graphs.GraphedStep) they are carved from ONE arena that a single
# memset node clears at the head of the graph.  GraphedStep sizes the arena with an eager dry run of the step.
# ------------------------------------------------------------------------------------------------
_ARENA = {"mode": None, "need": 0, "buf": None, "off": 0}


def arena_measure():
    _ARENA.update(mode="measure", need=0, buf=None, off=0)


def arena_begin(buf):
    """buf: uint8 tensor the caller has cleared (on the capturing stream, inside the graph)"""
    _ARENA.update(mode="carve", buf=buf, off=0)


def arena_end():
    need = _ARENA["need"]
    _ARENA.update(mode=None, buf=None, off=0, need=0)
    return need


def zeros(shape, dtype, device):
    """torch.zeros(shape), or a slice of the captured step's cleared arena"""
    mode = _ARENA["mode"]
    if mode is None:
        return torch.zeros(shape, dtype=dtype, device=device)
    n = 1
    for v in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)):
        n *= int(v)
    nbytes = (n * torch.empty(0, dtype=dtype).element_size() + 255) // 256 * 256
    if mode == "measure":
        _ARENA["need"] += nbytes
        return torch.zeros(shape, dtype=dtype, device=device)
    buf, off = _ARENA["buf"], _ARENA["off"]
    if buf is None or buf.device != torch.device(device) or off + nbytes > buf.numel():
        return torch.zeros(shape, dtype=dtype, device=device)     # (a shape the dry run did not see)
    _ARENA["off"] = off + nbytes
    return buf[off:off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(shape)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mpnn_b200: CUDA tensors required (no CPU fallback); got %s" % t.device)


# ------------------------------------------------------------------------------------------------
# edge-network trunk  (reference edge_network.py:14-21 minus the last Linear)
# ------------------------------------------------------------------------------------------------
class EdgeTrunkFn(torch.autograd.Function):
    """x = trunk(rows): rows [R, ef] -> x [R, PP] (PP = P rounded up to 4, pad columns are zero)."""

    @staticmethod
    def forward(ctx, rows, w_tied, n_tied, *growth):
        lib = _lib.load()
        _need_cuda(rows, w_tied)
        rows = f32c(rows)
        w_tied = f32c(w_tied)
        G = len(growth) // 2
        gw = [f32c(t) for t in growth[:G]]
        gb = [f32c(t) for t in growth[G:]]
        R, ef = rows.shape
        P = w_tied.shape[0]
        x_off = ctypes.c_longlong(0)
        ldx = ctypes.c_int(0)
        total = lib.mpnn_edge_trunk_saved_floats(R, ef, G, P, n_tied, ctypes.byref(x_off), ctypes.byref(ldx))
        if total < 0:
            raise RuntimeError("mpnn_b200: inconsistent edge_map layer plan (ef=%d, growth=%d, P=%d)" % (ef, G, P))
        saved = torch.empty(total, dtype=torch.float32, device=rows.device)
        ws = workspace(lib.mpnn_edge_trunk_workspace_bytes(R, ef, G, P), rows.device)
        check(lib.mpnn_edge_trunk_fwd(ptr(rows), R, ef, G, ptr_array(gw), ptr_array(gb), ptr(w_tied), P, n_tied,
                                      ptr(saved), ptr(ws), ws.numel(), stream()), "edge_trunk_fwd")
        PP = ldx.value
        x = saved[x_off.value:x_off.value + R * PP].view(R, PP)
        ctx.save_for_backward(rows, w_tied, saved, *gw)
        ctx.dims = (R, ef, G, P, n_tied, PP)
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, dx):
        lib = _lib.load()
        rows, w_tied, saved = ctx.saved_tensors[:3]
        gw = list(ctx.saved_tensors[3:])
        R, ef, G, P, n_tied, PP = ctx.dims
        if not (dx.dim() == 2 and dx.stride(1) == 1 and dx.stride(0) >= P and dx.dtype == torch.float32):
            dx = dx.contiguous().float()
        lddx = dx.stride(0)
        dev = rows.device
        d_w_tied = torch.empty_like(w_tied)
        d_gw = [torch.empty_like(w) for w in gw]
        d_gb = [torch.empty(w.shape[0], dtype=torch.float32, device=dev) for w in gw]
        d_rows = torch.empty_like(rows) if ctx.needs_input_grad[0] else None
        ws = workspace(lib.mpnn_edge_trunk_workspace_bytes(R, ef, G, P), dev)
        check(lib.mpnn_edge_trunk_bwd(ptr(rows), R, ef, G, ptr_array(gw), ptr(w_tied), P, n_tied, ptr(saved),
                                      ctypes.c_void_p(dx.data_ptr()), lddx, ptr_array(d_gw), ptr_array(d_gb),
                                      ptr(d_w_tied), ptr(d_rows), ptr(ws), ws.numel(), stream()), "edge_trunk_bwd")
        return (d_rows, d_w_tied, None) + tuple(d_gw) + tuple(d_gb)


# ------------------------------------------------------------------------------------------------
# gather-sum over an index list (CSR one way, CSC the other) -- deterministic, scatter-free both ways
# ------------------------------------------------------------------------------------------------
class GatherSumFn(torch.autograd.Function):
    """out[r,:] = sum_{k in ptr[r]:ptr[r+1]} src[idx[k],:]  (idx None -> k);  backward uses the transposed lists."""

    @staticmethod
    def forward(ctx, src, ptr_t, idx, n_out, t_ptr, t_idx):
        lib = _lib.load()
        _need_cuda(src)
        src = f32c(src)
        width = src.shape[1]
        out = torch.empty(n_out, width, dtype=torch.float32, device=src.device)
        check(lib.mpnn_segment_sum(ptr(src), ptr(ptr_t), ptr(idx), n_out, width, width, ptr(out), width, 0, 1.0,
                                   stream()), "segment_sum")
        ctx.t = (t_ptr, t_idx, src.shape[0], width)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        t_ptr, t_idx, n_src, width = ctx.t
        dout = f32c(dout)
        dsrc = torch.empty(n_src, width, dtype=torch.float32, device=dout.device)
        check(lib.mpnn_segment_sum(ptr(dout), ptr(t_ptr), ptr(t_idx), n_src, width, width, ptr(dsrc), width, 0, 1.0,
                                   stream()), "segment_sum")
        return dsrc, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# fused message + aggregation
# ------------------------------------------------------------------------------------------------
class EdgeMessageFn(torch.autograd.Function):
    """M[i] = W~ . ( sum_{e in E(i)} alpha_e x~_e (x) g_e + x~_0 (x) Q_i ) (+ beta)   -- see csrc/message.cu.

    X [E+1, ldx] trunk output (last row = x_0); G either node states H [n_rows, nf] (gather=True: g_e = H[src_e])
    or explicit per-edge vectors [E, nf]; alpha [E] or None; Q [n_rows, nf] or None; W_last [mf*nf, P],
    B_last [mf*nf] (reference edge_map[-1]), beta [mf] or None.
    """

    @staticmethod
    def forward(ctx, X, G, alpha, Q, W_last, B_last, beta, el, gather, nf, mf, P):
        lib = _lib.load()
        _need_cuda(X, G, W_last)
        assert X.stride(1) == 1
        ldx = X.stride(0)
        G = f32c(G)
        alpha_c = f32c(alpha) if alpha is not None else None
        Q_c = f32c(Q) if Q is not None else None
        W_last, B_last = f32c(W_last), f32c(B_last)
        beta_c = f32c(beta) if beta is not None else None
        dev = X.device
        n_wt = lib.mpnn_message_wt_floats(nf, mf, P)
        if n_wt < 0:
            raise RuntimeError("mpnn_b200: node/message feature width > 128 is not supported yet")
        Wt = torch.empty(n_wt, dtype=torch.float32, device=dev)
        check(lib.mpnn_message_prepare(ptr(W_last), ptr(B_last), nf, mf, P, ptr(Wt), stream()), "message_prepare")
        M = torch.empty(el.n_rows, mf, dtype=torch.float32, device=dev)
        gidx = el.edge_src if gather else None
        check(lib.mpnn_message_fwd(ptr(el.row_ptr), ptr(el.edge_dst), ptr(gidx), None, ptr(alpha_c),
                                   ctypes.c_void_p(X.data_ptr()), ldx, el.E, ptr(G), G.stride(0), ptr(Q_c), ptr(Wt),
                                   ptr(beta_c), el.n_rows, nf, mf, P, ptr(M), stream()), "message_fwd")
        ctx.save_for_backward(X, G, alpha_c, Q_c, Wt, beta_c)
        ctx.meta = (el, gather, nf, mf, P, ldx)
        return M

    @staticmethod
    @once_differentiable
    def backward(ctx, dM):
        lib = _lib.load()
        X, G, alpha, Q, Wt, beta = ctx.saved_tensors
        el, gather, nf, mf, P, ldx = ctx.meta
        dev = X.device
        dM = f32c(dM)
        E = el.E
        ldt = (P + 1 + 3) // 4 * 4
        T = torch.zeros(E + 1, ldt, dtype=torch.float32, device=dev)
        dG = torch.zeros(max(E, 1), nf, dtype=torch.float32, device=dev)
        dQ = torch.empty_like(Q) if Q is not None else None
        dalpha = torch.empty(max(E, 1), dtype=torch.float32, device=dev) if alpha is not None else None
        dW = torch.empty(mf * nf, P, dtype=torch.float32, device=dev)
        dB = torch.empty(mf * nf, dtype=torch.float32, device=dev)
        dbeta = torch.empty(mf, dtype=torch.float32, device=dev) if beta is not None else None
        ws = workspace(lib.mpnn_message_bwd_workspace_bytes(el.n_rows, nf, mf, P), dev)
        gidx = el.edge_src if gather else None
        check(lib.mpnn_message_bwd(ptr(el.row_ptr), ptr(el.edge_dst), ptr(gidx), None, ptr(alpha),
                                   ctypes.c_void_p(X.data_ptr()), ldx, E, ptr(G), G.stride(0), ptr(Q), ptr(Wt),
                                   ptr(beta), el.n_rows, E, nf, mf, P, ptr(dM), ptr(T), ldt, ptr(dG), ptr(dQ),
                                   ptr(dalpha), ptr(dW), ptr(dB), ptr(dbeta), ptr(ws), ws.numel(), stream()),
              "message_bwd")
        dX = T[:, :X.shape[1]]
        if gather:
            dGsrc = torch.empty(el.n_rows, nf, dtype=torch.float32, device=dev)
            check(lib.mpnn_segment_sum(ptr(dG), ptr(el.col_ptr), ptr(el.csc_eid), el.n_rows, nf, nf, ptr(dGsrc), nf, 0,
                                       1.0, stream()), "segment_sum")
        else:
            dGsrc = dG[:E]
        return (dX, dGsrc, dalpha[:E] if dalpha is not None else None, dQ, dW, dB, dbeta,
                None, None, None, None, None)


# ------------------------------------------------------------------------------------------------
# GRU update
# ------------------------------------------------------------------------------------------------
class SharedGradSession(object):
    """One forward pass through a GRU cell that is applied at every message-passing step (basic_model.py:50-58).  The
    steps' backward calls leave their per-CTA weight-gradient partials in one slab (`GRUFn.backward`); the hub node
    (`GRUParamHubFn`), which autograd runs after all of them, reduces the slab ONCE.  Replaces T reductions and the
    4 (T-1) accumulate kernels autograd would launch for the shared parameters."""

    def __init__(self, key):
        self.key, self.done, self.uses, self.filled = key, False, 0, 0
        self.slab, self.meta, self.handles = None, None, None


class GRUParamHubFn(torch.autograd.Function):
    """identity on (W_ih, W_hh, b_ih, b_hh); its backward turns the session's slab into the parameter gradients"""

    @staticmethod
    def forward(ctx, session, W_ih, W_hh, b_ih, b_hh):
        ctx.session = session
        ctx.params = (W_ih, W_hh, b_ih, b_hh)
        ctx.set_materialize_grads(False)
        return W_ih.detach(), W_hh.detach(), b_ih.detach(), b_hh.detach()

    @staticmethod
    @once_differentiable
    def backward(ctx, g_ih, g_hh, gb_ih, gb_hh):
        s = ctx.session
        s.done = True
        direct = [g_ih, g_hh, gb_ih, gb_hh]    # from steps that could not use the slab (returned real gradients)
        if not s.filled:
            return (None,) + tuple(direct)
        lib = _lib.load()
        rows, d, dev = s.meta
        # parameter gradients only: off the main stream when autograd will ASSIGN them (see EdgeNetTableFn.backward)
        side = (_side_ok() and all(g is None for g in direct)
                and all(getattr(p, "grad", None) is None for p in ctx.params))
        with (_on_side_stream(dev, [s.slab], lane=4) if side else _inline()):
            dW_ih = torch.empty(d, 3 * d, dtype=torch.float32, device=dev)
            dW_hh = torch.empty(d, 3 * d, dtype=torch.float32, device=dev)
            db_ih = torch.empty(3 * d, dtype=torch.float32, device=dev)
            db_hh = torch.empty(3 * d, dtype=torch.float32, device=dev)
            check(lib.mpnn_gru_bwd_params(ptr(s.slab), s.filled, rows, d, ptr(dW_ih), ptr(dW_hh), ptr(db_ih),
                                          ptr(db_hh), stream()), "gru_bwd_params")
        s.filled = 0                            # a second backward through a retained graph refills the slab
        out = [dW_ih, dW_hh, db_ih, db_hh]
        out = [o if g is None else o + g for o, g in zip(out, direct)]
        return (None,) + tuple(out)


class GRUFn(torch.autograd.Function):
    """reference gru_update.py:26-35,66-68 on flat [rows, d] tensors; mask [rows].  `session`: see SharedGradSession
    (None = parameter gradients are returned by every call)."""

    @staticmethod
    def forward(ctx, m, h, mask, W_ih, W_hh, b_ih, b_hh, session=None):
        ctx.session = session
        if session is not None:
            session.uses += 1
        lib = _lib.load()
        _need_cuda(m, h, mask, W_ih)
        m, h, mask = f32c(m), f32c(h), f32c(mask)
        W_ih, W_hh, b_ih, b_hh = f32c(W_ih), f32c(W_hh), f32c(b_ih), f32c(b_hh)
        rows, d = h.shape
        out = torch.empty_like(h)
        gates = torch.empty(rows, 4 * d, dtype=torch.float32, device=h.device)
        ws = workspace(lib.mpnn_gru_workspace_bytes(rows, d), h.device)
        pend = getattr(m, "_mpnn_pending_sum", None)
        if pend is not None:
            # TypedMessageTCFn left the aggregation to this kernel: its operand producer sums the per-edge messages of
            # every row while staging and writes the sums into m (saved for the backward below)
            del m._mpnn_pending_sum
            Y, row_ptr = pend
            check(lib.mpnn_gru_fwd_agg(ptr(Y), ptr(row_ptr), ptr(h), ptr(mask), ptr(W_ih), ptr(W_hh), ptr(b_ih),
                                       ptr(b_hh), rows, d, ptr(m), ptr(out), ptr(gates), ptr(ws), ws.numel(), stream()),
                  "gru_fwd_agg")
        else:
            check(lib.mpnn_gru_fwd(ptr(m), ptr(h), ptr(mask), ptr(W_ih), ptr(W_hh), ptr(b_ih), ptr(b_hh), rows, d,
                                   ptr(out), ptr(gates), ptr(ws), ws.numel(), stream()), "gru_fwd")
        ctx.save_for_backward(m, h, mask, W_ih, W_hh, gates)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        m, h, mask, W_ih, W_hh, gates = ctx.saved_tensors
        rows, d = h.shape
        dout = f32c(dout)
        dev = h.device
        dm, dh = torch.empty_like(m), torch.empty_like(h)
        s = ctx.session
        # the slab is only reduced by the hub node, which autograd reaches only when the cell's parameters are
        # differentiated in THIS backward pass (not e.g. torch.autograd.grad(loss, [afm])): otherwise return direct gradients
        if s is not None and not all(ctx.needs_input_grad[3:7]):
            s = None
        if s is not None:
            pb = lib.mpnn_gru_bwd_partial_bytes(rows, d)
            if pb and s.slab is None:
                s.slab = torch.empty(max(s.uses, 1) * pb, dtype=torch.uint8, device=dev)
                s.meta = (rows, d, dev)
            if pb and s.meta == (rows, d, dev) and (s.filled + 1) * pb <= s.slab.numel():
                check(lib.mpnn_gru_bwd_data(ptr(m), ptr(h), ptr(mask), ptr(W_ih), ptr(W_hh), ptr(gates), ptr(dout), rows,
                                            d, ptr(dm), ptr(dh), s.slab.data_ptr() + s.filled * pb, stream()),
                      "gru_bwd_data")
                s.filled += 1
                _note_produced(dm)
                return dm, dh, None, None, None, None, None, None
        dW_ih, dW_hh = torch.empty_like(W_ih), torch.empty_like(W_hh)
        db_ih = torch.empty(3 * d, dtype=torch.float32, device=dev)
        db_hh = torch.empty(3 * d, dtype=torch.float32, device=dev)
        ws = workspace(lib.mpnn_gru_workspace_bytes(rows, d), dev)
        check(lib.mpnn_gru_bwd(ptr(m), ptr(h), ptr(mask), ptr(W_ih), ptr(W_hh), ptr(gates), ptr(dout), rows, d, ptr(dm),
                               ptr(dh), ptr(dW_ih), ptr(dW_hh), ptr(db_ih), ptr(db_hh), ptr(ws), ws.numel(), stream()),
              "gru_bwd")
        _note_produced(dm)
        return dm, dh, None, dW_ih, dW_hh, db_ih, db_hh, None


# ------------------------------------------------------------------------------------------------
# masked batch norms
# ------------------------------------------------------------------------------------------------
class MaskBNFn(torch.autograd.Function):
    """reference mask_batch_norm.py:9-15 on x [rows, C], mask [rows]."""

    @staticmethod
    def forward(ctx, x, mask, eps):
        lib = _lib.load()
        _need_cuda(x, mask)
        x, mask = f32c(x), f32c(mask)
        rows, C = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        ws = _lib.clean_workspace(lib.mpnn_bn_workspace_bytes(rows, C), x.device)
        check(lib.mpnn_mask_bn_fwd(ptr(x), ptr(mask), rows, C, float(eps), ptr(y), ptr(stats), ptr(ws), ws.numel(),
                                   stream()), "mask_bn_fwd")
        ctx.save_for_backward(x, mask, stats)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x, mask, stats = ctx.saved_tensors
        rows, C = x.shape
        dy = f32c(dy)
        dx = torch.empty_like(x)
        ws = _lib.clean_workspace(lib.mpnn_bn_workspace_bytes(rows, C), x.device)
        check(lib.mpnn_mask_bn_bwd(ptr(x), ptr(mask), ptr(dy), ptr(stats), rows, C, ptr(dx), ptr(ws), ws.numel(),
                                   stream()), "mask_bn_bwd")
        return dx, None, None


class TypedRowBNFn(torch.autograd.Function):
    """Masked batch norm of a categorical bond tensor in row space (csrc/bn.cu k_row_bn_*): rows [R, F] distinct rows,
    a [R] their mask (adjacency) values, cnt [R] their multiplicities (data, not differentiated).  `bn1d` selects
    MaskBatchNorm1d's arithmetic (masked mean, eps outside the root, affine, running statistics: mask_batch_norm.py:20-38),
    else MaskBatchNorm's (:9-15)."""

    @staticmethod
    def forward(ctx, rows, a, cnt, gamma, beta, running_mean, running_var, bn1d, training, momentum, eps):
        lib = _lib.load()
        _need_cuda(rows, a, cnt)
        rows, a, cnt = f32c(rows), f32c(a), f32c(cnt)
        gamma_c = f32c(gamma) if gamma is not None else None
        beta_c = f32c(beta) if beta is not None else None
        R, F = rows.shape
        y = torch.empty_like(rows)
        stats = torch.empty(3 * F + 1, dtype=torch.float32, device=rows.device)
        check(lib.mpnn_row_bn_fwd(ptr(rows), ptr(a), ptr(cnt), R, F, ptr(gamma_c), ptr(beta_c), ptr(running_mean),
                                  ptr(running_var), int(bn1d), int(not bn1d), int(training), float(momentum or 0.0),
                                  float(eps), ptr(y), ptr(stats), stream()), "row_bn_fwd")
        ctx.save_for_backward(rows, a, cnt, gamma_c, stats)
        ctx.cfg = (int(bn1d), int(training), beta is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        rows, a, cnt, gamma, stats = ctx.saved_tensors
        bn1d, training, has_beta = ctx.cfg
        R, F = rows.shape
        dy = f32c(dy)
        dev = rows.device
        dx = torch.empty_like(rows) if ctx.needs_input_grad[0] else None
        dgamma = torch.empty(F, dtype=torch.float32, device=dev) if gamma is not None else None
        dbeta = torch.empty(F, dtype=torch.float32, device=dev) if has_beta else None
        check(lib.mpnn_row_bn_bwd(ptr(rows), ptr(a), ptr(cnt), R, F, ptr(gamma), ptr(stats), ptr(dy), bn1d, int(not bn1d),
                                  training, ptr(dx), ptr(dgamma), ptr(dbeta), stream()), "row_bn_bwd")
        return dx, None, None, dgamma, dbeta, None, None, None, None, None, None


class MaskBN1dFn(torch.autograd.Function):
    """reference mask_batch_norm.py:20-38; running buffers are updated in place when training."""

    @staticmethod
    def forward(ctx, x, mask, weight, bias, running_mean, running_var, training, momentum, eps):
        lib = _lib.load()
        _need_cuda(x, mask)
        x, mask = f32c(x), f32c(mask)
        weight_c = f32c(weight) if weight is not None else None
        bias_c = f32c(bias) if bias is not None else None
        rows, C = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        ws = _lib.clean_workspace(lib.mpnn_bn_workspace_bytes(rows, C), x.device)
        check(lib.mpnn_mask_bn1d_fwd(ptr(x), ptr(mask), ptr(weight_c), ptr(bias_c), ptr(running_mean), ptr(running_var),
                                     rows, C, int(training), float(momentum), float(eps), ptr(y), ptr(stats), ptr(ws),
                                     ws.numel(), stream()), "mask_bn1d_fwd")
        if training:
            ctx.save_for_backward(x, mask, weight_c, stats)
        else:  # eval normalises with the running statistics as they are now
            ctx.save_for_backward(x, mask, weight_c, running_mean.clone(), running_var.clone())
        ctx.cfg = (bool(training), float(eps))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        training, eps = ctx.cfg
        if training:
            x, mask, weight, stats = ctx.saved_tensors
            rm = rv = None
        else:
            x, mask, weight, rm, rv = ctx.saved_tensors
            stats = None
        rows, C = x.shape
        dy = f32c(dy)
        dev = x.device
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=torch.float32, device=dev)
        db = torch.empty(C, dtype=torch.float32, device=dev)
        ws = _lib.clean_workspace(lib.mpnn_bn_workspace_bytes(rows, C), dev)
        check(lib.mpnn_mask_bn1d_bwd(ptr(x), ptr(mask), ptr(dy), ptr(weight), ptr(stats), ptr(rm), ptr(rv), rows, C,
                                     int(training), eps, ptr(dx), ptr(dw), ptr(db), ptr(ws), ws.numel(), stream()),
              "mask_bn1d_bwd")
        return dx, None, (dw if weight is not None else None), (db if weight is not None else None), None, None, \
            None, None, None


# ------------------------------------------------------------------------------------------------
# graph-level readout
# ------------------------------------------------------------------------------------------------
class GraphLevelOutputFn(torch.autograd.Function):
    """reference graph_level_output.py:30-47; x [B,N,F2], mask [B,N,1] or None."""

    @staticmethod
    def forward(ctx, x, mask, Wi, bi, Wj, bj):
        lib = _lib.load()
        _need_cuda(x, Wi)
        params = (Wi, bi, Wj, bj)
        x = f32c(x)
        mask_c = f32c(mask) if mask is not None else None
        Wi, bi, Wj, bj = f32c(Wi), f32c(bi), f32c(Wj), f32c(bj)
        B, N, F2 = x.shape
        O = Wi.shape[0]
        dev = x.device
        out = torch.empty(B, O, dtype=torch.float32, device=dev)
        u = torch.empty(B * N, O, dtype=torch.float32, device=dev)
        v = torch.empty(B * N, O, dtype=torch.float32, device=dev)
        UV = torch.empty(B, 2, O, dtype=torch.float32, device=dev) if mask is None else None
        ws = workspace(lib.mpnn_glo_workspace_bytes(B, N, F2, O), dev)
        check(lib.mpnn_glo_fwd(ptr(x), ptr(mask_c), ptr(Wi), ptr(bi), ptr(Wj), ptr(bj), B, N, F2, O, ptr(out), ptr(u),
                               ptr(v), ptr(UV), ptr(ws), ws.numel(), stream()), "glo_fwd")
        ctx.save_for_backward(x, mask_c, Wi, Wj, u, v, UV)
        ctx.params = params
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        x, mask, Wi, Wj, u, v, UV = ctx.saved_tensors
        B, N, F2 = x.shape
        O = Wi.shape[0]
        dev = x.device
        dout = f32c(dout)
        dx = torch.empty_like(x)
        dWi, dWj = torch.empty_like(Wi), torch.empty_like(Wj)
        dbi = torch.empty(O, dtype=torch.float32, device=dev)
        dbj = torch.empty(O, dtype=torch.float32, device=dev)
        ws = workspace(lib.mpnn_glo_workspace_bytes(B, N, F2, O), dev)
        if lib.mpnn_glo_bwd_split_supported(int(mask is not None), F2, O):
            # dx continues the main backward chain; the reduction of the weight-gradient partials feeds parameter gradients
            # only: inside a captured step it is a side branch (when autograd will ASSIGN those gradients, see
            # EdgeNetTableFn.backward)
            check(lib.mpnn_glo_bwd_data(ptr(x), ptr(mask), ptr(Wi), ptr(Wj), ptr(u), ptr(v), ptr(dout), B, N, F2, O, ptr(dx),
                                        ptr(ws), ws.numel(), stream()), "glo_bwd_data")
            side = (_side_ok() and all(ctx.needs_input_grad[2:6])
                    and all(getattr(t, "grad", None) is None for t in ctx.params))
            with (_on_side_stream(dev, [ws], lane=7) if side else _inline()):
                check(lib.mpnn_glo_bwd_params(ptr(ws), B, N, F2, O, ptr(dWi), ptr(dbi), ptr(dWj), ptr(dbj), stream()),
                      "glo_bwd_params")
            return dx, None, dWi, dbi, dWj, dbj
        check(lib.mpnn_glo_bwd(ptr(x), ptr(mask), ptr(Wi), ptr(Wj), ptr(u), ptr(v), ptr(UV), ptr(dout), B, N, F2, O,
                               ptr(dx), ptr(dWi), ptr(dbi), ptr(dWj), ptr(dbj), ptr(ws), ws.numel(), stream()),
              "glo_bwd")
        return dx, None, dWi, dbi, dWj, dbj


# ------------------------------------------------------------------------------------------------
# row gather with a scatter-free backward (transposed index lists)
# ------------------------------------------------------------------------------------------------
class GatherRowsFn(torch.autograd.Function):
    """out[t,:] = src[idx[t],:];  backward: dsrc[s,:] = sum_{k in t_ptr[s]:t_ptr[s+1]} dout[t_idx[k] (or k),:]."""

    @staticmethod
    def forward(ctx, src, idx, t_ptr, t_idx):
        lib = _lib.load()
        _need_cuda(src)
        src = f32c(src)
        T, width = idx.shape[0], src.shape[1]
        out = torch.empty(T, width, dtype=torch.float32, device=src.device)
        if T:
            unit = torch.arange(T + 1, dtype=torch.int32, device=src.device)
            check(lib.mpnn_segment_sum(ptr(src), ptr(unit), ptr(idx), T, width, width, ptr(out), width, 0, 1.0,
                                       stream()), "segment_sum")
        ctx.t = (t_ptr, t_idx, src.shape[0], width)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        t_ptr, t_idx, n_src, width = ctx.t
        dout = f32c(dout)
        dsrc = torch.empty(n_src, width, dtype=torch.float32, device=dout.device)
        check(lib.mpnn_segment_sum(ptr(dout), ptr(t_ptr), ptr(t_idx), n_src, width, width, ptr(dsrc), width, 0, 1.0,
                                   stream()), "segment_sum")
        return dsrc, None, None, None


class TypeGatherFn(torch.autograd.Function):
    """out[e,:] = per_type[uid[e],:] for every edge slot; backward: the fixed-order sum over each type's edge list
    (type_ptr / type_eid of graph.TypedInfo; the rows behind the last type, e.g. the zero row, get zero)."""

    @staticmethod
    def forward(ctx, per_type, ti):
        lib = _lib.load()
        _need_cuda(per_type)
        per_type = f32c(per_type)
        T, width = ti.uid.shape[0], per_type.shape[1]
        out = torch.empty(T, width, dtype=torch.float32, device=per_type.device)
        if T:
            unit = torch.arange(T + 1, dtype=torch.int32, device=per_type.device)
            check(lib.mpnn_segment_sum(ptr(per_type), ptr(unit), ptr(ti.uid), T, width, width, ptr(out), width, 0, 1.0,
                                       stream()), "segment_sum")
        ctx.t = (ti, per_type.shape[0], width)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        ti, n_src, width = ctx.t
        dout = f32c(dout)
        ti.wait_sorted()
        d = zeros((n_src, width), torch.float32, dout.device)
        n_types = min(n_src, ti.type_ptr.shape[0] - 1)
        check(lib.mpnn_segment_sum(ptr(dout), ptr(ti.type_ptr), ptr(ti.type_eid), n_types, width, width, ptr(d), width,
                                   0, 1.0, stream()), "segment_sum")
        return d, None


# ------------------------------------------------------------------------------------------------
# linear layer on flat rows (plumbing GEMM of the library; y = x W^T + b, W is an nn.Linear weight)
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        lib = _lib.load()
        _need_cuda(x, W)
        x, W = f32c(x), f32c(W)
        b_c = f32c(b) if b is not None else None
        R, K = x.shape
        O = W.shape[0]
        y = torch.empty(R, O, dtype=torch.float32, device=x.device)
        check(lib.mpnn_gemm(ptr(x), ptr(W), ptr(y), R, O, K, K, 1, 1, K, O, ptr(b_c), 0, None, 0, stream()), "gemm")
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x, W = ctx.saved_tensors
        dy = f32c(dy)
        R, K = x.shape
        O = W.shape[0]
        dev = x.device
        dx = torch.empty_like(x)
        dW = torch.empty_like(W)
        check(lib.mpnn_gemm(ptr(dy), ptr(W), ptr(dx), R, K, O, O, 1, K, 1, K, None, 0, None, 0, stream()), "gemm")
        ws = workspace(lib.mpnn_gemm_workspace_bytes(O, K, R) + lib.mpnn_colsum_workspace_bytes(R, O), dev)
        check(lib.mpnn_gemm(ptr(dy), ptr(x), ptr(dW), O, K, R, 1, O, K, 1, K, None, 0, ptr(ws), ws.numel(), stream()),
              "gemm")
        db = None
        if ctx.has_bias:
            db = torch.empty(O, dtype=torch.float32, device=dev)
            check(lib.mpnn_colsum(ptr(dy), None, R, O, O, 0, ptr(db), 0, ptr(ws), ws.numel(), stream()), "colsum")
        return dx, dW, db


# ------------------------------------------------------------------------------------------------
# softmax(logits) * V   (attention gate of AttEdgeNetwork; row softmax of WAdjMsgAgg with V=None)
# ------------------------------------------------------------------------------------------------
class SoftmaxMulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, V):
        lib = _lib.load()
        _need_cuda(logits)
        logits = f32c(logits)
        V_c = f32c(V) if V is not None else None
        rows, n = logits.shape
        gate = torch.empty_like(logits)
        out = torch.empty_like(logits)
        check(lib.mpnn_softmax_mul_fwd(ptr(logits), ptr(V_c), rows, n, ptr(gate), ptr(out), stream()), "softmax_mul_fwd")
        ctx.save_for_backward(gate, V_c)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        gate, V = ctx.saved_tensors
        dout = f32c(dout)
        rows, n = gate.shape
        dlogits = torch.empty_like(gate)
        dV = torch.empty_like(gate) if V is not None else None
        check(lib.mpnn_softmax_mul_bwd(ptr(gate), ptr(V), ptr(dout), rows, n, ptr(dlogits), ptr(dV), stream()),
              "softmax_mul_bwd")
        return dlogits, dV


# ------------------------------------------------------------------------------------------------
# dense weighted aggregation over senders (stand-alone aggregator contract on [B,N,N,mf] messages)
# ------------------------------------------------------------------------------------------------
class DenseAggFn(torch.autograd.Function):
    """out[b,i,:] = sum_j w[b,i,j] * messages[b,i,j,:]"""

    @staticmethod
    def forward(ctx, messages, w):
        lib = _lib.load()
        _need_cuda(messages, w)
        messages, w = f32c(messages), f32c(w)
        B, N, N2, mf = messages.shape
        out = torch.empty(B, N, mf, dtype=torch.float32, device=messages.device)
        check(lib.mpnn_dense_agg_fwd(ptr(messages), ptr(w), B * N, N2, mf, ptr(out), stream()), "dense_agg_fwd")
        ctx.save_for_backward(messages, w)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        messages, w = ctx.saved_tensors
        B, N, N2, mf = messages.shape
        dout = f32c(dout)
        dm = torch.empty_like(messages) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        check(lib.mpnn_dense_agg_bwd(ptr(messages), ptr(w), ptr(dout), B * N, N2, mf, ptr(dm), ptr(dw), stream()),
              "dense_agg_bwd")
        return dm, dw


# ------------------------------------------------------------------------------------------------
# Set2Vec readout
# ------------------------------------------------------------------------------------------------
_S2V_PERSISTENT = [True]


def set2vec_persistent(enabled):
    """Set2Vec's loop in the persistent kernels (csrc/s2v_persist.cu; default) or as per-iteration launches; returns the
    previous setting.  (Part of the op-graph keys: a captured per-shape graph holds one of the two forms.)"""
    prev = _S2V_PERSISTENT[0]
    _lib.load().mpnn_set2vec_set_persistent(1 if enabled else 0)
    _S2V_PERSISTENT[0] = bool(enabled)
    return prev


class Set2VecFn(torch.autograd.Function):
    """reference set2vec.py:93-151 ("default" inner product).  Wcat [2F,4F] = [w_hi|w_hf|w_hg|w_ho], bcat [4F];
    m0 [B,2F] / c0 [B,F]: caller-supplied initial state (set2vec.py:111-117), None = zeros."""

    @staticmethod
    def forward(ctx, X, mask, Wcat, bcat, Wq, we, steps, m0=None, c0=None):
        lib = _lib.load()
        _need_cuda(X, Wcat)
        X = f32c(X)
        mask_c = f32c(mask) if mask is not None else None
        Wcat, bcat, Wq, we = f32c(Wcat), f32c(bcat), f32c(Wq), f32c(we)
        m0_c = f32c(m0) if m0 is not None else None
        c0_c = f32c(c0) if c0 is not None else None
        B, N, F = X.shape
        dev = X.device

        def make_bufs():
            return [torch.empty(B, 2 * F, dtype=torch.float32, device=dev),
                    torch.empty(lib.mpnn_set2vec_saved_floats(B, N, F, steps), dtype=torch.float32, device=dev),
                    workspace(lib.mpnn_set2vec_workspace_bytes(B, N, F), dev)]

        def run(ins, bufs):
            x, mk, wc, bc, wq, w_e, m_0, c_0 = ins
            out, saved, ws = bufs
            check(lib.mpnn_set2vec_fwd(ptr(x), ptr(mk), ptr(wc), ptr(bc), ptr(wq), ptr(w_e), ptr(m_0), ptr(c_0), B, N, F,
                                       steps, ptr(out), ptr(saved), ptr(ws), ws.numel(), stream()), "set2vec_fwd")

        out, saved = _graphed(("set2vec_fwd", B, N, F, steps, mask is not None, m0 is not None, c0 is not None,
                               dev.index, _S2V_PERSISTENT[0]), run, [X, mask_c, Wcat, bcat, Wq, we, m0_c, c0_c], make_bufs, 2)
        ctx.save_for_backward(X, mask_c, Wcat, Wq, we, saved, m0_c, c0_c)
        ctx.steps = steps
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        X, mask, Wcat, Wq, we, saved, m0, c0 = ctx.saved_tensors
        B, N, F = X.shape
        dev = X.device
        dout = f32c(dout)
        steps = ctx.steps

        def make_bufs():
            return [torch.empty_like(X), torch.empty_like(Wcat), torch.empty(4 * F, dtype=torch.float32, device=dev),
                    torch.empty_like(Wq), torch.empty_like(we),
                    torch.empty(B, 2 * F, dtype=torch.float32, device=dev) if m0 is not None else None,
                    torch.empty(B, F, dtype=torch.float32, device=dev) if c0 is not None else None,
                    workspace(lib.mpnn_set2vec_bwd_workspace_bytes(B, N, F, steps), dev)]

        def run(ins, bufs):
            x, mk, wc, wq, w_e, m_0, c_0, sv, do = ins
            dX, dWcat, dbcat, dWq, dwe, dm0, dc0, ws = bufs
            check(lib.mpnn_set2vec_bwd(ptr(x), ptr(mk), ptr(wc), ptr(wq), ptr(w_e), ptr(m_0), ptr(c_0), ptr(sv), ptr(do),
                                       B, N, F, steps, ptr(dX), ptr(dWcat), ptr(dbcat), ptr(dWq), ptr(dwe), ptr(dm0),
                                       ptr(dc0), ptr(ws), ws.numel(), stream()), "set2vec_bwd")

        if m0 is None and c0 is None:
            dX, dWcat, dbcat, dWq, dwe, dm0, dc0 = _graphed(
                ("set2vec_bwd", B, N, F, steps, mask is not None, dev.index, _S2V_PERSISTENT[0]), run,
                [X, mask, Wcat, Wq, we, None, None, saved, dout], make_bufs, 7)
        else:
            bufs = make_bufs()
            run([X, mask, Wcat, Wq, we, m0, c0, saved, dout], bufs)
            dX, dWcat, dbcat, dWq, dwe, dm0, dc0 = bufs[:7]
        return dX, None, dWcat, dbcat, dWq, dwe, None, dm0, dc0


class LSTMCellHiddenFn(torch.autograd.Function):
    """reference set2vec.py:68-75 (the input-less LSTM cell on its own): (hprev [B,2F], cprev [B,F], Wcat, bcat) ->
    (h' [B,F], c' [B,F])."""

    @staticmethod
    def forward(ctx, hprev, cprev, Wcat, bcat):
        lib = _lib.load()
        _need_cuda(hprev, cprev, Wcat)
        hprev, cprev, Wcat, bcat = f32c(hprev), f32c(cprev), f32c(Wcat), f32c(bcat)
        B, K = hprev.shape
        F = Wcat.shape[1] // 4
        dev = hprev.device
        pre = torch.empty(B, 4 * F, dtype=torch.float32, device=dev)
        check(lib.mpnn_gemm(ptr(hprev), ptr(Wcat), ptr(pre), B, 4 * F, K, K, 1, 4 * F, 1, 4 * F, ptr(bcat), 0, None, 0,
                            stream()), "gemm")
        gates = torch.empty(B, 4 * F, dtype=torch.float32, device=dev)
        c = torch.empty(B, F, dtype=torch.float32, device=dev)
        tc = torch.empty(B, F, dtype=torch.float32, device=dev)
        h = torch.empty(B, F, dtype=torch.float32, device=dev)
        check(lib.mpnn_lstm_hidden_fwd(ptr(pre), ptr(cprev), B, F, ptr(gates), ptr(c), ptr(tc), ptr(h), stream()),
              "lstm_hidden_fwd")
        ctx.save_for_backward(hprev, cprev, Wcat, gates, tc)
        return h, c

    @staticmethod
    @once_differentiable
    def backward(ctx, dh, dc):
        lib = _lib.load()
        hprev, cprev, Wcat, gates, tc = ctx.saved_tensors
        B, K = hprev.shape
        F = Wcat.shape[1] // 4
        dev = hprev.device
        dh = f32c(dh) if dh is not None else torch.zeros(B, F, dtype=torch.float32, device=dev)
        dc = f32c(dc) if dc is not None else torch.zeros(B, F, dtype=torch.float32, device=dev)
        dpre = torch.empty(B, 4 * F, dtype=torch.float32, device=dev)
        dcprev = torch.empty(B, F, dtype=torch.float32, device=dev)
        check(lib.mpnn_lstm_hidden_bwd(ptr(gates), ptr(tc), ptr(cprev), ptr(dh), ptr(dc), B, F, ptr(dpre), ptr(dcprev),
                                       stream()), "lstm_hidden_bwd")
        dhprev = torch.empty_like(hprev)
        check(lib.mpnn_gemm(ptr(dpre), ptr(Wcat), ptr(dhprev), B, K, 4 * F, 4 * F, 1, 1, 4 * F, K, None, 0, None, 0,
                            stream()), "gemm")
        dW = torch.empty_like(Wcat)
        ws = workspace(lib.mpnn_gemm_workspace_bytes(K, 4 * F, B) + lib.mpnn_colsum_workspace_bytes(B, 4 * F), dev)
        check(lib.mpnn_gemm(ptr(hprev), ptr(dpre), ptr(dW), K, 4 * F, B, 1, K, 4 * F, 1, 4 * F, None, 0, ptr(ws),
                            ws.numel(), stream()), "gemm")
        db = torch.empty(4 * F, dtype=torch.float32, device=dev)
        check(lib.mpnn_colsum(ptr(dpre), None, B, 4 * F, 4 * F, 0, ptr(db), 0, ptr(ws), ws.numel(), stream()), "colsum")
        return dhprev, dcprev, dW, db


# ------------------------------------------------------------------------------------------------
# typed path: edge network on the DISTINCT bond rows -> table of matrices; gather message kernel
# (csrc/dedup.cu, csrc/typed.cu)
# ------------------------------------------------------------------------------------------------
def typed_dp(nf, mf):
    return _lib.load().mpnn_typed_dp(nf, mf)


def tc_dp(nf, mf):
    """padded width (64/128/256) when the tcgen05 typed path (csrc/tc_message.cu) serves the shape, else -1"""
    return _lib.load().mpnn_tc_dp(int(nf), int(mf))


def table_dp(nf, mf):
    """padded width of the table of per-type matrices, whichever typed kernel family serves the shape (-1: none)"""
    d = typed_dp(nf, mf)
    return d if d >= 0 else tc_dp(nf, mf)


_ENET_LANES = (0, 2, 3)
_ENET_LANE = 0

# forward applications of a parameter set (keyed by its first tensor) that have not been back-propagated yet
_PARAM_USES = {}


def _note_param_use(t):
    key = (t.data_ptr(), t._version)
    _PARAM_USES[key] = _PARAM_USES.get(key, 0) + 1
    if len(_PARAM_USES) > 256:      # forward passes that never reach a backward
        _PARAM_USES.clear()
    return key


def _single_param_use(key):
    """True when the backward being run is the only pending application of the parameter set (consumes the use)"""
    n = _PARAM_USES.pop(key, 1)
    if n == 1:
        return True
    rem = abs(n) - 1            # several applications: none of their backward passes may leave the main stream
    if rem > 0:
        _PARAM_USES[key] = -rem
    return False


class EdgeNetTableFn(torch.autograd.Function):
    """Fused growth layers + 50 tied layers + last Linear on the distinct rows (P <= 64):
    urows [R, ef] -> table T[u][l][k], tableT T[u][k][l]  ([R, DP, DP] each; reference edge_network.py:14-21,37)."""

    @staticmethod
    def forward(ctx, urows, w_tied, n_tied, W_last, B_last, nf, mf, *growth):
        lib = _lib.load()
        _need_cuda(urows, w_tied, W_last)
        urows, w_tied, W_last, B_last = f32c(urows), f32c(w_tied), f32c(W_last), f32c(B_last)
        G = len(growth) // 2
        gw = [f32c(t) for t in growth[:G]]
        gb = [f32c(t) for t in growth[G:]]
        R, ef = urows.shape
        P = w_tied.shape[0]
        DP = table_dp(nf, mf)
        dev = urows.device
        saved = torch.empty(lib.mpnn_enet_saved_floats(R, G, n_tied), dtype=torch.float32, device=dev)
        table = torch.empty(R, DP, DP, dtype=torch.float32, device=dev)
        tableT = torch.empty(R, DP, DP, dtype=torch.float32, device=dev)
        check(lib.mpnn_enet_fwd(ptr(urows), R, ef, G, ptr_array(gw), ptr_array(gb), ptr(w_tied), P, n_tied, ptr(W_last),
                                ptr(B_last), nf, mf, ptr(saved), ptr(table), ptr(tableT), stream()), "enet_fwd")
        ctx.save_for_backward(urows, w_tied, W_last, saved, *gw)
        ctx.dims = (R, ef, G, P, n_tied, nf, mf)
        ctx.mark_non_differentiable(tableT)
        ctx.use_key = _note_param_use(w_tied)
        return table, tableT

    @staticmethod
    @once_differentiable
    def backward(ctx, dT, _dTt):
        lib = _lib.load()
        urows, w_tied, W_last, saved = ctx.saved_tensors[:4]
        gw = list(ctx.saved_tensors[4:])
        R, ef, G, P, n_tied, nf, mf = ctx.dims
        dev = urows.device
        dT = f32c(dT) if dT is not None else zeros((R, table_dp(nf, mf), table_dp(nf, mf)), torch.float32, dev)
        d_w_tied = torch.empty_like(w_tied)
        d_W_last = torch.empty_like(W_last)
        d_B_last = torch.empty(W_last.shape[0], dtype=torch.float32, device=dev)
        d_gw = [torch.empty_like(w) for w in gw]
        d_gb = [torch.empty(w.shape[0], dtype=torch.float32, device=dev) for w in gw]
        d_rows = torch.empty_like(urows) if ctx.needs_input_grad[0] else None
        # only parameter gradients come out of this call (d_rows is needed upstream: stay on the main stream then)
        # deferred join is only safe when autograd ASSIGNS these gradients (param.grad is None: no kernel touches
        # them before the join); if it has to accumulate into an existing .grad the work stays on the main stream
        # ... and when this is the ONLY application of these weights since their last backward: with two producers
        # autograd sums the two gradients on the main stream, which does not wait for the side lanes
        side = (_side_ok() and d_rows is None and _single_param_use(ctx.use_key)
                and all(getattr(t, "grad", None) is None for t in [w_tied, W_last] + gw))
        # the edge networks of a model are independent of each other: their backward chains (57 CTAs each) go to
        # different lanes round-robin, so the last one does not queue behind the others
        global _ENET_LANE
        _ENET_LANE = (_ENET_LANE + 1) % len(_ENET_LANES)
        lane = _ENET_LANES[_ENET_LANE]
        ready = _ready_get(dT)
        if not side and ready is not None:
            torch.cuda.current_stream(dev).wait_event(ready)   # dT was produced on a side stream
        with (_on_side_stream(dev, [dT, urows, w_tied, W_last, saved] + gw, lane=lane, after=ready) if side
              else _inline()):
            ws = workspace(lib.mpnn_enet_workspace_bytes(R, ef, G, P), dev)
            check(lib.mpnn_enet_bwd(ptr(urows), R, ef, G, ptr_array(gw), ptr(w_tied), P, n_tied, ptr(W_last), nf, mf,
                                    ptr(saved), ptr(dT), ptr_array(d_gw), ptr_array(d_gb), ptr(d_w_tied),
                                    ptr(d_W_last), ptr(d_B_last), ptr(d_rows), ptr(ws), ws.numel(), stream()),
                  "enet_bwd")
        return (d_rows, d_w_tied, None, d_W_last, d_B_last, None, None) + tuple(d_gw) + tuple(d_gb)


class TableLayoutFn(torch.autograd.Function):
    """flat [R, mf*nf] (last Linear's output on the distinct rows) -> table, tableT [R, DP, DP] (any trunk width)."""

    @staticmethod
    def forward(ctx, flat, nf, mf):
        lib = _lib.load()
        _need_cuda(flat)
        flat = f32c(flat)
        R = flat.shape[0]
        DP = table_dp(nf, mf)
        table = torch.empty(R, DP, DP, dtype=torch.float32, device=flat.device)
        tableT = torch.empty(R, DP, DP, dtype=torch.float32, device=flat.device)
        check(lib.mpnn_table_from_flat(ptr(flat), R, nf, mf, ptr(table), ptr(tableT), stream()), "table_from_flat")
        ctx.dims = (R, nf, mf)
        ctx.mark_non_differentiable(tableT)
        ctx.set_materialize_grads(False)
        return table, tableT

    @staticmethod
    @once_differentiable
    def backward(ctx, dT, _dTt):
        lib = _lib.load()
        R, nf, mf = ctx.dims
        if dT is None:
            return None, None, None
        dT = f32c(dT)
        dflat = torch.empty(R, mf * nf, dtype=torch.float32, device=dT.device)
        check(lib.mpnn_table_to_flat(ptr(dT), R, nf, mf, ptr(dflat), stream()), "table_to_flat")
        return dflat, None, None


class MultiEdgeNetTableFn(torch.autograd.Function):
    """EdgeNetTableFn for K sibling networks (one EdgeNetwork per message-passing step, normed_basic_model.py:24-27: same
    layer plan, same distinct rows, different weights) as ONE launch each way (csrc/typed.cu k_enet_*_multi).
    params: per network [w_tied, W_last, B_last, growth weights (G), growth biases (G)].
    Returns (table_0, tableT_0, table_1, tableT_1, ...)."""

    @staticmethod
    def forward(ctx, urows, n_tied, nf, mf, K, G, *params):
        lib = _lib.load()
        per = 3 + 2 * G
        assert len(params) == K * per
        urows = f32c(urows)
        _need_cuda(urows)
        nets = [[f32c(t) for t in params[k * per:(k + 1) * per]] for k in range(K)]
        R, ef = urows.shape
        P = nets[0][0].shape[0]
        DP = table_dp(nf, mf)
        dev = urows.device
        nsaved = lib.mpnn_enet_saved_floats(R, G, n_tied)
        saved = [torch.empty(nsaved, dtype=torch.float32, device=dev) for _ in range(K)]
        tables = [torch.empty(R, DP, DP, dtype=torch.float32, device=dev) for _ in range(K)]
        tablesT = [torch.empty(R, DP, DP, dtype=torch.float32, device=dev) for _ in range(K)]
        gw = [w for n in nets for w in n[3:3 + G]]
        gb = [b for n in nets for b in n[3 + G:3 + 2 * G]]
        check(lib.mpnn_enet_fwd_multi(K, ptr(urows), R, ef, G, ptr_array(gw), ptr_array(gb),
                                      ptr_array([n[0] for n in nets]), P, n_tied, ptr_array([n[1] for n in nets]),
                                      ptr_array([n[2] for n in nets]), nf, mf, ptr_array(saved), ptr_array(tables),
                                      ptr_array(tablesT), stream()), "enet_fwd_multi")
        ctx.save_for_backward(urows, *([t for n in nets for t in n] + saved))
        ctx.dims = (R, ef, G, P, n_tied, nf, mf, K)
        ctx.use_keys = [_note_param_use(n[0]) for n in nets]
        out = []
        for k in range(K):
            out += [tables[k], tablesT[k]]
        ctx.mark_non_differentiable(*tablesT)
        ctx.set_materialize_grads(False)     # (else autograd fills a zero tensor per transposed table, every step)
        return tuple(out)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        lib = _lib.load()
        R, ef, G, P, n_tied, nf, mf, K = ctx.dims
        per = 3 + 2 * G
        urows = ctx.saved_tensors[0]
        flat = ctx.saved_tensors[1:1 + K * per]
        saved = list(ctx.saved_tensors[1 + K * per:])
        nets = [list(flat[k * per:(k + 1) * per]) for k in range(K)]
        dev = urows.device
        DP = table_dp(nf, mf)
        dTs, ready = [], None
        for k in range(K):
            g = grads[2 * k]
            if g is None:
                g = zeros((R, DP, DP), torch.float32, dev)
            else:
                ev = _ready_get(g)
                ready = ev if ev is not None else ready
                g = f32c(g)
            dTs.append(g)
        d_nets = [[torch.empty_like(t) for t in n] for n in nets]
        need_rows = ctx.needs_input_grad[0]
        d_rows = [torch.empty_like(urows) for _ in range(K)] if need_rows else None
        single = all([_single_param_use(k) for k in ctx.use_keys])
        side = (_side_ok() and not need_rows and single
                and all(getattr(t, "grad", None) is None for n in nets for t in n))
        if not side and ready is not None:
            torch.cuda.current_stream(dev).wait_event(ready)
        with (_on_side_stream(dev, dTs + [urows] + list(flat) + saved, lane=0, after=ready) if side else _inline()):
            ws = workspace(K * lib.mpnn_enet_workspace_bytes(R, ef, G, P), dev)
            gw = [w for n in nets for w in n[3:3 + G]]
            d_gw = [w for n in d_nets for w in n[3:3 + G]]
            d_gb = [b for n in d_nets for b in n[3 + G:3 + 2 * G]]
            check(lib.mpnn_enet_bwd_multi(K, ptr(urows), R, ef, G, ptr_array(gw), ptr_array([n[0] for n in nets]), P,
                                          n_tied, ptr_array([n[1] for n in nets]), nf, mf, ptr_array(saved),
                                          ptr_array(dTs), ptr_array(d_gw), ptr_array(d_gb),
                                          ptr_array([n[0] for n in d_nets]), ptr_array([n[1] for n in d_nets]),
                                          ptr_array([n[2] for n in d_nets]), ptr_array(d_rows) if d_rows else None,
                                          ptr(ws), ws.numel(), stream()), "enet_bwd_multi")
        d_urows = None
        if need_rows:
            d_urows = d_rows[0]
            for k in range(1, K):
                d_urows = d_urows + d_rows[k]
        return (d_urows, None, None, None, None, None) + tuple(t for n in d_nets for t in n)


class TableHolder(object):
    """Rides on a table tensor produced by EdgeNetTableFn: how many message functions consume it in this forward pass.
    With exactly one consumer its gradient dT goes straight from TypedMessageFn.backward into EdgeNetTableFn.backward
    (no autograd accumulation in between), so the table half of the message backward can run on the side stream too."""

    def __init__(self):
        self.uses = 0


class TypedMessageFn(torch.autograd.Function):
    """M[i] = sum_{e in E(i)} alpha_e T[uid_e]^T H[src_e]  (+ HEAD terms: zero-row matrix on all non-bonded pairs,
    + beta).  H [n_rows, nf]; table/tableT from EdgeNetTableFn / TableLayoutFn; alpha [E] or None (not
    differentiated: the adjacency value); head selects edge_network.py:50-51 over edge_network.py:52."""

    @staticmethod
    def forward(ctx, H, table, tableT, beta, el, alpha, head, nf, mf, holder=None):
        lib = _lib.load()
        _need_cuda(H, table)
        ctx.holder = holder
        if holder is not None:
            holder.uses += 1
        H, table, tableT = f32c(H), f32c(table), f32c(tableT)
        beta_c = f32c(beta) if beta is not None else None
        alpha_c = f32c(alpha) if alpha is not None else None
        ti = el.typed()
        dev = H.device
        S = None
        if head:
            S = torch.empty(el.B, nf, dtype=torch.float32, device=dev)
            check(lib.mpnn_graph_sum(ptr(H), el.B, el.N, nf, ptr(S), stream()), "graph_sum")
        M = torch.empty(el.n_rows, mf, dtype=torch.float32, device=dev)
        check(lib.mpnn_tmsg_fwd(ptr(el.row_ptr), ptr(el.edge_src), ptr(ti.uid), ptr(alpha_c), ptr(H), ptr(table), ptr(S),
                                ptr(beta_c), el.n_rows, el.N, nf, mf, ti.zero_type, ptr(M), stream()), "tmsg_fwd")
        ctx.save_for_backward(H, table, tableT, S, alpha_c)
        ctx.meta = (el, nf, mf, beta is not None)
        return M

    @staticmethod
    @once_differentiable
    def backward(ctx, dM):
        lib = _lib.load()
        H, table, tableT, S, alpha = ctx.saved_tensors
        el, nf, mf, has_beta = ctx.meta
        ti = el.typed()
        dev = H.device
        dM = f32c(dM)
        need_dH, need_dT = ctx.needs_input_grad[0], ctx.needs_input_grad[1]

        def run(dH, dT):
            if dT is not None:
                ti.wait_sorted()
            ws = workspace(lib.mpnn_tmsg_bwd_workspace_bytes(el.Ecap, ti.Ucap, nf, mf, el.B), dev)
            check(lib.mpnn_tmsg_bwd(ptr(el.row_ptr), ptr(el.col_ptr), ptr(el.csc_eid), ptr(el.edge_src),
                                    ptr(el.edge_dst), ptr(ti.uid), ptr(ti.type_ptr), ptr(ti.type_eid), ptr(ti.counts),
                                    ptr(alpha), ptr(H), ptr(table), ptr(tableT), ptr(S), el.n_rows, H.shape[0], el.B,
                                    el.N, nf, mf, el.Ecap, ti.Ucap, ptr(dM), ptr(dH), ptr(dT), ptr(ws), ws.numel(),
                                    stream()), "tmsg_bwd")

        # The sender-state gradient continues the main backward chain (and is skipped when the senders are data: the
        # reference's loops pass the INPUT features to every message function, normed_basic_model.py:58).  The table
        # gradient only feeds the edge network's parameter gradients: with a single consumer of the table it runs on the
        # side stream, where EdgeNetTableFn.backward picks it up in stream order.
        dH = dT = None
        h = ctx.holder
        side = _side_ok() and need_dT and h is not None and h.uses == 1
        if need_dH:
            dH = torch.empty_like(H)
            if not side and need_dT:
                dT = torch.empty_like(table)
            run(dH, dT)
        if need_dT and dT is None:
            # fork behind the kernel that produced dM (the GRU backward registers it), not behind whatever the main
            # stream has been given since
            cm = (_on_side_stream(dev, [dM, H, table, tableT, S, alpha], lane=1, after=_ready_get(dM)) if side
                  else _inline())
            with cm:
                dT = torch.empty_like(table)
                run(None, dT)
            if side:
                _ready_put(dT, cm.done)
        dbeta = None
        if has_beta:
            dbeta = torch.empty(mf, dtype=torch.float32, device=dev)
            ws2 = workspace(lib.mpnn_colsum_workspace_bytes(el.n_rows, mf), dev)
            check(lib.mpnn_colsum(ptr(dM), None, el.n_rows, mf, mf, 0, ptr(dbeta), 0, ptr(ws2), ws2.numel(), stream()),
                  "colsum")
        return dH, dT, None, dbeta, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# the whole T-step loop as one persistent kernel each way (csrc/chain.cu)
# ------------------------------------------------------------------------------------------------
CHAIN_ENABLED = os.environ.get("MPNN_B200_FUSED_CHAIN", "1") != "0"


def chain_supported(d, T):
    return bool(CHAIN_ENABLED and _lib.load().mpnn_chain_supported(int(d), int(T)))


AFTER_CHAIN_FWD = []   # one-shot callbacks run right behind the launch of the fused step kernel's forward


_REAL_ROWS = {}     # (mask identity) -> (mask, list, event): the rows with mask != 0, computed once per batch


def real_rows(mask, side=False, fork=None):
    """(list [rows+1] int32 or None, event or None) for the step kernels: the rows of `mask` that are real, so every CTA
    owns the same number of them.  With side=True the (tiny, single-block) kernel is enqueued on a side stream, forked
    from the current position of the main stream: `modules._fused_chain` calls it before the compaction and the edge
    networks, which it then overlaps.  Batches beyond mpnn_real_rows_max() rows use the kernels' round-robin deal."""
    lib = _lib.load()
    m = mask.reshape(-1)
    rows = m.shape[0]
    if rows > lib.mpnn_real_rows_max() or not m.is_cuda:
        return None, None
    key = (m.data_ptr(), m._version, rows, m.device.index)
    hit = _REAL_ROWS.get(key)
    if hit is not None:
        return hit[1], hit[2]
    m = f32c(m)
    dev = m.device
    lst = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    ws = torch.empty(2 * (rows + 1), dtype=torch.int32, device=dev)
    ev = None
    if side and SIDE_STREAM_ENABLED:
        main = torch.cuda.current_stream(dev)
        _, s_ = _side_stream(dev, lane=6)
        if fork is None:       # (else: an event the caller recorded earlier on the main stream)
            fork = torch.cuda.Event()
            fork.record(main)
        s_.wait_event(fork)
        with torch.cuda.stream(s_):
            check(lib.mpnn_real_rows(ptr(m), rows, ptr(lst), ptr(ws), ws.numel() * 4, stream()), "real_rows")
            ev = torch.cuda.Event()
            ev.record(s_)
        for t_ in (m, lst, ws):
            t_.record_stream(s_)
        _note_forward_side_work(dev, lane=6)
    else:
        check(lib.mpnn_real_rows(ptr(m), rows, ptr(lst), ptr(ws), ws.numel() * 4, stream()), "real_rows")
    _REAL_ROWS.clear()
    _REAL_ROWS[key] = (mask, lst, ev)
    return lst, ev


class ChainFn(torch.autograd.Function):
    """h_T of  h <- bn_t(GRU(sum_e alpha_e T_t[uid_e]^T H0[src_e], h) * mask), t = 0..T-1  (see csrc/chain.cu).

    args: H0 [rows,d], h_init [rows,d], mask [rows], el, bn (list of T dicts: kind 0/1/2, module, training, eps,
    momentum), W_ih, W_hh, b_ih, b_hh, then T tables, T transposed tables, then (gamma, beta) of every kind-2 step that
    has an affine transform."""

    @staticmethod
    def forward(ctx, H0, h_init, mask, el, bn, W_ih, W_hh, b_ih, b_hh, *rest):
        lib = _lib.load()
        _need_cuda(H0, h_init, mask, W_ih)
        real = real_rows(mask)      # (list tensor or None, event or None): enqueued early by modules._fused_chain
        T = len(bn)
        tables = [f32c(t) for t in rest[:T]]
        tablesT = [f32c(t) for t in rest[T:2 * T]]
        affine = list(rest[2 * T:])
        H0, h_init, mask = f32c(H0), f32c(h_init), f32c(mask)
        W_ih, W_hh, b_ih, b_hh = f32c(W_ih), f32c(W_hh), f32c(b_ih), f32c(b_hh)
        rows, d = h_init.shape
        dev = h_init.device
        ti = el.typed()
        kinds = (ctypes.c_int * T)(*[b["kind"] for b in bn])
        training = (ctypes.c_int * T)(*[int(b["training"]) for b in bn])
        eps = (ctypes.c_float * T)(*[float(b["eps"]) for b in bn])
        mom = (ctypes.c_float * T)(*[float(b["momentum"] or 0.0) for b in bn])
        ptrs, ai, aff_idx = [], 0, []
        for b in bn:
            g = be = None
            if b["kind"] == 2 and b["affine"]:
                g, be = f32c(affine[ai]), f32c(affine[ai + 1])
                aff_idx.append(ai)
                ai += 2
            else:
                aff_idx.append(-1)
            ptrs += [g, be, b.get("running_mean"), b.get("running_var")]
        keep = [p for p in ptrs if p is not None]
        bn_ptrs = ptr_array(ptrs)
        out = zeros(h_init.shape, torch.float32, h_init.device)   # rows with mask == 0 are skipped by the kernel: exact zeros
        saved = torch.empty(lib.mpnn_chain_saved_floats(rows, d, T), dtype=torch.float32, device=dev)
        if real[1] is not None:
            torch.cuda.current_stream(dev).wait_event(real[1])
        ws = _lib.clean_workspace(lib.mpnn_chain_workspace_bytes(rows, d, T), dev, "chain")
        alpha = f32c(el.edge_w) if el.edge_w is not None else None
        args = (ptr(el.row_ptr), ptr(el.edge_src), ptr(ti.uid), ptr(alpha), el.Ecap, ti.zero_type, ptr(H0), ptr(h_init),
                ptr(mask), ptr_array(tables), T, ptr(W_ih), ptr(W_hh), ptr(b_ih), ptr(b_hh), kinds, training, eps, mom,
                bn_ptrs, rows, d, ptr(real[0]), ptr(saved))
        check(lib.mpnn_chain_fwd(*(args + (ptr(out), ptr(ws), ws.numel(), stream()))), "chain_fwd")
        while AFTER_CHAIN_FWD:      # one-shot callbacks (graphs.GraphedStep: fork the next batch's preprocessing here)
            AFTER_CHAIN_FWD.pop(0)()
        ctx.save_for_backward(H0, h_init, mask, W_ih, W_hh, b_ih, b_hh, saved, alpha, *(tables + tablesT + affine))
        ctx.meta = (el, bn, T, rows, d, aff_idx, keep, real[0])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        el, bn, T, rows, d, aff_idx, keep, real_list = ctx.meta
        H0, h_init, mask, W_ih, W_hh, b_ih, b_hh, saved, alpha = ctx.saved_tensors[:9]
        rest = ctx.saved_tensors[9:]
        tables, tablesT, affine = list(rest[:T]), list(rest[T:2 * T]), list(rest[2 * T:])
        dev = h_init.device
        ti = el.typed()
        # the reference models concatenate the final state with afm (normed_basic_model.py:59): the gradient arrives as a
        # column slice of a wider array and the kernel reads it in place through its row stride
        if (dout.dtype == torch.float32 and dout.dim() == 2 and dout.shape[0] > 1 and dout.stride(1) == 1
                and dout.stride(0) >= d):
            dout_ld = int(dout.stride(0))
        else:
            dout = f32c(dout)
            dout_ld = d
        need_H0, need_h = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        kinds = (ctypes.c_int * T)(*[b["kind"] for b in bn])
        training = (ctypes.c_int * T)(*[int(b["training"]) for b in bn])
        eps = (ctypes.c_float * T)(*[float(b["eps"]) for b in bn])
        mom = (ctypes.c_float * T)(*[float(b["momentum"] or 0.0) for b in bn])
        ptrs = []
        for t, b in enumerate(bn):
            g = be = None
            if aff_idx[t] >= 0:
                g, be = affine[aff_idx[t]], affine[aff_idx[t] + 1]
            ptrs += [g, be, b.get("running_mean"), b.get("running_var")]
        # padded rows (mask == 0) are skipped by the kernel: their gradients are exact zeros
        dM = zeros((T, rows, d), torch.float32, dev)
        dh = zeros(h_init.shape, torch.float32, dev) if need_h else None
        dW_ih, dW_hh = torch.empty_like(W_ih), torch.empty_like(W_hh)
        db_ih, db_hh = torch.empty_like(b_ih), torch.empty_like(b_hh)
        d_aff = [torch.empty_like(a) for a in affine]
        gptrs = []
        for t in range(T):
            gptrs += ([d_aff[aff_idx[t]], d_aff[aff_idx[t] + 1]] if aff_idx[t] >= 0 else [None, None])
        ws = _lib.clean_workspace(lib.mpnn_chain_workspace_bytes(rows, d, T), dev, "chain")
        check(lib.mpnn_chain_bwd(ptr(el.row_ptr), ptr(el.edge_src), ptr(ti.uid), ptr(alpha), el.Ecap, ti.zero_type, ptr(H0),
                                 ptr(h_init), ptr(mask), ptr_array(tables), T, ptr(W_ih), ptr(W_hh), ptr(b_ih), ptr(b_hh),
                                 kinds, training, eps, mom, ptr_array(ptrs), rows, d, ptr(real_list), ptr(saved),
                                 ctypes.c_void_p(dout.data_ptr()), dout_ld, ptr(dM),   # (row-strided: see above)
                                 ptr(dh), ptr(dW_ih), ptr(dW_hh), ptr(db_ih), ptr(db_hh), ptr_array(gptrs), ptr(ws),
                                 ws.numel(), stream()), "chain_bwd")
        # ---- message gradients -> table gradients (parameter-only: side lane) and sender gradients -------------------
        # steps that share a table (basic_model.py:57: one EdgeNetwork, reuse_graph_tensors) share H0 too, so their
        # message gradients add up before the (linear) table / sender gradient kernels
        groups = {}
        for t in range(T):
            groups.setdefault(tables[t].data_ptr(), []).append(t)
        need_T = [ctx.needs_input_grad[9 + t] for t in range(T)]
        dTs = [None] * T
        dH0 = None
        produced = torch.cuda.Event() if _side_ok() else None
        if produced is not None:
            produced.record(torch.cuda.current_stream(dev))

        def run(dMsum, t0, dH, dT):
            if dT is not None:
                ti.wait_sorted()
            wsb = workspace(lib.mpnn_tmsg_bwd_workspace_bytes(el.Ecap, ti.Ucap, d, d, el.B), dev)
            check(lib.mpnn_tmsg_bwd(ptr(el.row_ptr), ptr(el.col_ptr), ptr(el.csc_eid), ptr(el.edge_src),
                                    ptr(el.edge_dst), ptr(ti.uid), ptr(ti.type_ptr), ptr(ti.type_eid), ptr(ti.counts),
                                    ptr(alpha), ptr(H0), ptr(tables[t0]), ptr(tablesT[t0]), None, el.n_rows, H0.shape[0],
                                    el.B, el.N, d, d, el.Ecap, ti.Ucap, ptr(dMsum), ptr(dH), ptr(dT), ptr(wsb),
                                    wsb.numel(), stream()), "tmsg_bwd")

        glist = list(groups.values())
        if need_H0:
            for ts in glist:
                t0 = ts[0]
                dMsum = dM[t0] if len(ts) == 1 else dM[ts].sum(0)
                dHt = torch.empty_like(H0)
                run(dMsum, t0, dHt, None)
                dH0 = dHt if dH0 is None else dH0 + dHt
        want = [ts for ts in glist if any(need_T[t] for t in ts)]
        if want:
            # the table gradients of all steps as ONE launch pair; they only feed the edge networks' parameter
            # gradients: side lane, behind the step kernel
            K = len(want)
            single = all(len(ts) == 1 for ts in want)
            if single and [ts[0] for ts in want] == list(range(T)):
                dMs = dM
            else:
                dMs = torch.stack([dM[ts[0]] if len(ts) == 1 else dM[ts].sum(0) for ts in want])
            side = _side_ok()
            DPt = tables[0].shape[-1]
            cm = _on_side_stream(dev, [dMs, H0, alpha], lane=1, after=produced if dMs is dM else None) if side \
                else _inline()
            with cm:
                ti.wait_sorted()
                dTall = torch.empty(K, ti.Ucap + 1, DPt, DPt, dtype=torch.float32, device=dev)
                wsb = workspace(K * lib.mpnn_tmsg_bwd_workspace_bytes(el.Ecap, ti.Ucap, d, d, 1), dev)
                check(lib.mpnn_tmsg_bwd_table_multi(K, ptr(el.edge_src), ptr(el.edge_dst), ptr(ti.uid), ptr(ti.type_ptr),
                                                    ptr(ti.type_eid), ptr(ti.counts), ptr(alpha), ptr(H0), el.n_rows, d,
                                                    d, el.Ecap, ti.Ucap, ptr(dMs), ptr(dTall), ptr(wsb), wsb.numel(),
                                                    stream()), "tmsg_bwd_table_multi")
            for k, ts in enumerate(want):
                dT = dTall[k]
                if side:
                    _ready_put(dT, cm.done)
                dTs[ts[0]] = dT   # the other steps of the group point at the same table: autograd sums, they get None
        return (dH0, dh, None, None, None, dW_ih, dW_hh, db_ih, db_hh) + tuple(dTs) + (None,) * T + tuple(d_aff)


class BilinearEdgeFn(torch.autograd.Function):
    """reference bilinear_edge_network.py:25-37 on the compacted pairs: Y[e, p] = h[src_e]^T X_e[:, p, :] h[dst_e] with
    X_e the bond row viewed [nf, nf, nf].  H [n_rows, nf], X [E(+1), nf^3] -> Y [E, nf]."""

    @staticmethod
    def forward(ctx, H, X, el):
        lib = _lib.load()
        _need_cuda(H, X)
        H, X = f32c(H), f32c(X)
        nf = H.shape[1]
        E = el.E
        Y = torch.empty(max(E, 1), nf, dtype=torch.float32, device=H.device)
        check(lib.mpnn_bilinear_fwd(ptr(el.edge_src), ptr(el.edge_dst), ptr(X), X.shape[1], ptr(H), E, nf, ptr(Y),
                                    stream()), "bilinear_fwd")
        ctx.save_for_backward(H, X)
        ctx.el = el
        return Y[:E]

    @staticmethod
    @once_differentiable
    def backward(ctx, dY):
        lib = _lib.load()
        H, X = ctx.saved_tensors
        el = ctx.el
        nf = H.shape[1]
        dY = f32c(dY)
        dH = torch.empty_like(H) if ctx.needs_input_grad[0] else None
        dX = torch.zeros_like(X) if ctx.needs_input_grad[1] else None   # the trailing all-zero row gets no gradient
        check(lib.mpnn_bilinear_bwd(ptr(el.row_ptr), ptr(el.col_ptr), ptr(el.csc_eid), ptr(el.edge_src),
                                    ptr(el.edge_dst), ptr(X), X.shape[1], ptr(H), ptr(dY), el.n_rows, el.E, nf, ptr(dH),
                                    ptr(dX), X.shape[1], stream()), "bilinear_bwd")
        return dH, dX, None


class TypedMessageTCFn(torch.autograd.Function):
    """Tensor-core form of the typed message path for feature widths 33..256 (csrc/tc_message.cu):
    M[i] = sum_{e in E(i)} alpha_e T[uid_e]^T H[src_e] as a grouped TF32 GEMM over the type-sorted edge tiles
    (tcgen05.mma, accumulator in TMEM) + the fixed-order CSR segmented sum.  Backward: the same GEMM kernel on the
    gathered message gradients + CSC segmented sum, and dT[u] = sum_e alpha_e H[src_e] (x) dM[dst_e] with K = edges.
    alpha is the edge list's own weight (adj value) when weighted is True, else 1.
    The HEAD-form extras (edge_network.py:50-51) are composed around this op in modules.EdgeNetwork."""

    @staticmethod
    def forward(ctx, H, table, tableT, el, weighted, nf, mf, defer_sum=False):
        """defer_sum: the caller hands M straight to GRUFn (modules._wide_chain), whose tensor-core kernel does the CSR sum
        in its operand producer and fills M; M is NOT valid before that call."""
        lib = _lib.load()
        _need_cuda(H, table)
        H, table, tableT = f32c(H), f32c(table), f32c(tableT)
        ti = el.typed()
        DP = table.shape[-1]
        dev = H.device
        plan = ti.tc_plan(el)
        Y = torch.empty(max(el.Ecap, 1), mf, dtype=torch.float32, device=dev)
        ws = workspace(lib.mpnn_tc_edge_gemm_workspace_bytes(ti.Ucap, DP), dev)
        check(lib.mpnn_tc_edge_gemm(ptr(plan), el.Ecap, ti.Ucap, ptr(ti.type_eid), 0, ptr(H), nf, nf, ptr(tableT), DP,
                                    1 if weighted else 0, ptr(Y), mf, mf, ptr(ws), ws.numel(), stream()), "tc_edge_gemm")
        M = torch.empty(el.n_rows, mf, dtype=torch.float32, device=dev)
        if defer_sum and lib.mpnn_gru_agg_supported(mf):
            M._mpnn_pending_sum = (Y, el.row_ptr)
        else:
            check(lib.mpnn_segment_sum(ptr(Y), ptr(el.row_ptr), None, el.n_rows, mf, mf, ptr(M), mf, 0, 1.0, stream()),
                  "segment_sum")
        ctx.save_for_backward(H, table)
        ctx.meta = (el, nf, mf, DP, bool(weighted))
        return M

    @staticmethod
    @once_differentiable
    def backward(ctx, dM):
        lib = _lib.load()
        H, table = ctx.saved_tensors
        el, nf, mf, DP, weighted = ctx.meta
        ti = el.typed()
        dev = H.device
        dM = f32c(dM)
        plan = ti.tc_plan(el)
        dH = dT = None
        if ctx.needs_input_grad[0]:
            dG = torch.empty(max(el.Ecap, 1), nf, dtype=torch.float32, device=dev)
            ws = workspace(lib.mpnn_tc_edge_gemm_workspace_bytes(ti.Ucap, DP), dev)
            check(lib.mpnn_tc_edge_gemm(ptr(plan), el.Ecap, ti.Ucap, ptr(ti.type_eid), 1, ptr(dM), mf, mf, ptr(table), DP,
                                        1 if weighted else 0, ptr(dG), nf, nf, ptr(ws), ws.numel(), stream()),
                  "tc_edge_gemm")
            dH = torch.empty_like(H)
            check(lib.mpnn_segment_sum(ptr(dG), ptr(el.col_ptr), ptr(el.csc_eid), el.n_rows, nf, nf, ptr(dH), nf, 0,
                                       1.0, stream()), "segment_sum")
        if ctx.needs_input_grad[1]:
            dT = torch.empty_like(table)
            ws = workspace(lib.mpnn_tc_table_grad_workspace_bytes(ti.Ucap, DP), dev)
            check(lib.mpnn_tc_table_grad(ptr(plan), el.Ecap, ti.Ucap, ptr(H), nf, ptr(dM), mf, DP, 1 if weighted else 0,
                                         ptr(dT), ptr(ws), ws.numel(), stream()), "tc_table_grad")
        return dH, dT, None, None, None, None, None, None
