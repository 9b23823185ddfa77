"""On-device edge list of a padded batch (CSR by receiver + CSC by sender), built once per batch.

Layout contract: the reference's padded tensors (collate_2d_graphs, pre_process/data_loader.py:50-70):
bfm [B,N,N,ef], adj [B,N,N].  The compacted form is what every message kernel consumes.
"""
import collections
import os

import torch

from . import _lib


class EdgeList(object):
    """Device arrays describing the non-zero atom pairs of one padded batch.

    row_ptr [B*N+1] int32, edge_dst/edge_src [E] int32 (flat node ids b*N+i / b*N+j), edge_w [E] (adj value),
    rows [E+1, ef] (bond rows; the LAST row is the all-zero row x_0), col_ptr [B*N+1], csc_eid [E].
    Edge order is row-major (b,i,j) == torch.nonzero order (bit-exact contract).
    """

    def __init__(self, B, N, ef, E, row_ptr, col_ptr, edge_src, edge_dst, edge_w, rows, csc_eid):
        self.B, self.N, self.ef, self.E = B, N, ef, E
        self.n_rows = B * N
        self.row_ptr, self.col_ptr = row_ptr, col_ptr
        self.edge_src, self.edge_dst, self.edge_w = edge_src, edge_dst, edge_w
        self.rows, self.csc_eid = rows, csc_eid
        self._csc_dst = None
        self._per_edge = None
        self._typed = None
        self.Ecap = E           # allocated edge slots (== E unless built with explicit capacities)

    def typed(self):
        """Distinct bond rows of the batch + edges grouped by distinct row (csrc/dedup.cu), built once."""
        if self._typed is None:
            self._typed = dedup_rows(self)
        return self._typed

    def per_edge_view(self):
        """The same edge list for message functions that bring ONE explicit sender vector per edge (AttEdgeNetwork's
        gated states, att_edge_network.py:26) instead of gathering node states: sender row of edge e is e."""
        if self._per_edge is None:
            self._per_edge = PerEdgeView(self)
        return self._per_edge

    def neutralise_tail(self):
        """capacity mode: edge slots behind the last real edge -> sender / receiver / bond type 0, weight 0 (in place;
        the CSR / CSC / type lists never reach them, only per-slot kernels do)"""
        if self.E is not None or getattr(self, "_tail_done", False):
            return
        self._tail_done = True
        tail = torch.arange(self.Ecap, device=self.row_ptr.device) >= self.row_ptr[-1]
        self.edge_w.masked_fill_(tail, 0.0)
        self.edge_src.masked_fill_(tail, 0)
        self.edge_dst.masked_fill_(tail, 0)
        self.typed().uid.masked_fill_(tail, 0)

    @property
    def csc_dst(self):
        """receiver row of every CSC entry (for scatter-free transposed reductions)"""
        if self.E is None:
            raise RuntimeError("mpnn_b200: only the typed message path is available in capacity (graph-capture) mode")
        if self._csc_dst is None:
            self._csc_dst = self.edge_dst[self.csc_eid.long()].contiguous() if self.E else self.edge_dst
        return self._csc_dst


class PerEdgeView(object):
    """EdgeList whose senders are the edges themselves (identity edge_src / col_ptr / csc_eid)."""

    def __init__(self, el):
        dev = el.row_ptr.device
        E = el.E
        if E is None:
            # capacity mode: one sender vector per edge SLOT.  The slots behind the batch's last edge get valid indices
            # and weight 0, so that whatever is computed for them is finite and their gradients are exactly zero.
            E = el.Ecap
            el.neutralise_tail()
        self.B, self.N, self.ef, self.E, self.Ecap, self.n_rows = el.B, el.N, el.ef, el.E, el.Ecap, el.n_rows
        self.row_ptr, self.edge_dst, self.edge_w, self.rows = el.row_ptr, el.edge_dst, el.edge_w, el.rows
        self.edge_src = torch.arange(max(E, 1), dtype=torch.int32, device=dev)[:E]
        self.col_ptr = torch.arange(E + 1, dtype=torch.int32, device=dev)
        self.csc_eid = self.edge_src
        self._base = el

    def typed(self):
        return self._base.typed()


# ---- capacity mode (CUDA-graph capture): array sizes come from the caller, nothing is read back ----------
_CAPACITY = None          # (edge_capacity, unique_capacity) or None
_CAPTURED_COUNTS = []     # counts tensors of the edge lists built in capacity mode (overflow flags)
STATS = {"E": 0, "U": 0, "n_real": 0}  # largest edge / distinct-row counts seen by the eager path (sizes the capacities)


class capacities(object):
    """`with capacities(ecap, ucap):` edge lists are built with fixed array sizes and NO device->host read, so
    the whole step can be captured into a CUDA graph and replayed on other batches of the same padded shape.
    Only the typed message path is available in this mode; an overflow is flagged in counts[2]."""

    def __init__(self, edge_capacity, unique_capacity):
        self.cap = (int(edge_capacity), int(unique_capacity))

    def __enter__(self):
        global _CAPACITY
        self.prev = _CAPACITY
        _CAPACITY = self.cap
        if self.prev is None:
            # a new capacity session: the overflow guards the optimizer consults (optim.FusedAdam.step) are those of the
            # edge lists built INSIDE it, never the flags an earlier session left behind
            del _CAPTURED_COUNTS[:]
        return self

    def __exit__(self, *exc):
        global _CAPACITY
        _CAPACITY = self.prev
        return False


FUSED_PREP = os.environ.get("MPNN_B200_FUSED_PREP", "1") != "0"


def _type_sort_on_side_lane(uid, counts, Ecap, Ucap, dev, out=None):
    """type_ptr / type_eid / type_pos (edges grouped by distinct row, stable): only the backward's table-gradient kernels
    and the tensor-core plan read them, so the three dependent launches run on a side stream (a parallel branch of the
    captured step).  Returns (type_ptr, type_eid, type_pos, done event)."""
    from . import functional
    lib = _lib.load()
    if out is not None:
        type_ptr, type_eid, type_pos = out
    else:
        type_ptr = torch.empty(Ucap + 1, dtype=torch.int32, device=dev)
        type_eid = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
        type_pos = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    main = torch.cuda.current_stream(dev)
    _, side = functional._side_stream(dev, lane=5)
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        ws2 = _lib.workspace(lib.mpnn_type_sort_workspace_bytes(Ecap, Ucap), dev)
        _lib.check(lib.mpnn_type_sort(_lib.ptr(uid), _lib.ptr(counts), Ecap, Ucap, _lib.ptr(type_ptr),
                                      _lib.ptr(type_eid), _lib.ptr(type_pos), _lib.ptr(ws2), ws2.numel(),
                                      _lib.stream()), "type_sort")
        done = torch.cuda.Event()
        done.record(side)
    for t in (uid, counts, type_ptr, type_eid, type_pos):
        t.record_stream(side)
    functional._note_forward_side_work(dev, lane=5)
    return type_ptr, type_eid, type_pos, done


def _note_edge_overflow(counts, e_true, Ecap):
    """counts[0] <- true edge count, counts[2] |= (true count > capacity)   (multi-launch capacity path)"""
    counts[0:1].copy_(e_true[0:1])
    counts[2:3].bitwise_or_((e_true[0:1] > Ecap).to(torch.int32))


def _slab_views(slab, sizes):
    """consecutive 256-byte aligned int32 views of `slab`"""
    out, off = [], 0
    for n in sizes:
        out.append(slab[off:off + n])
        off += (n + 63) // 64 * 64
    return out


def _prep_edges(bfm_c, adj_c, B, N, ef, Ecap, Ucap, slab=False):
    """capacity mode, small batches: compaction + de-duplication as ONE cooperative launch (csrc/prep.cu).
    slab=True: every output array (incl. the type-sorted lists and the counts) is a view of ONE int32 allocation
    (`el.slab`), so a whole edge list can be handed over with one device copy (graphs.GraphedStep(pipeline_prep=True))."""
    lib = _lib.load()
    dev = bfm_c.device
    n_rows = B * N
    i32 = dict(dtype=torch.int32, device=dev)
    from . import functional
    sort_out, slab_t = None, None
    if slab:
        e1 = max(Ecap, 1)
        sizes = [n_rows + 1, n_rows + 1, e1, e1, e1, e1, e1, (Ucap + 1) * ef, 4, Ucap + 1, e1, e1]
        slab_t = torch.empty(sum((n + 63) // 64 * 64 for n in sizes), **i32)
        (row_ptr, col_ptr, edge_src, edge_dst, csc_eid, uid, edge_w, urows, counts, type_ptr, type_eid,
         type_pos) = _slab_views(slab_t, sizes)
        edge_src, edge_dst, csc_eid = edge_src[:Ecap], edge_dst[:Ecap], csc_eid[:Ecap]
        edge_w = edge_w.view(torch.float32)[:Ecap]
        urows = urows.view(torch.float32).view(Ucap + 1, ef)
        counts.zero_()
        sort_out = (type_ptr, type_eid, type_pos)
    else:
        row_ptr = torch.empty(n_rows + 1, **i32)
        col_ptr = torch.empty(n_rows + 1, **i32)
        edge_src = torch.empty(Ecap, **i32)
        edge_dst = torch.empty(Ecap, **i32)
        csc_eid = torch.empty(Ecap, **i32)
        uid = torch.empty(max(Ecap, 1), **i32)
        edge_w = torch.empty(Ecap, dtype=torch.float32, device=dev)
        urows = torch.empty(Ucap + 1, ef, dtype=torch.float32, device=dev)
        counts = functional.zeros((4,), torch.int32, dev)
    ws = _lib.clean_workspace(lib.mpnn_prep_workspace_bytes(B, Ucap), dev, "prep")
    _lib.check(lib.mpnn_prep_edges(_lib.ptr(bfm_c), _lib.ptr(adj_c), B, N, ef, Ecap, Ucap, _lib.ptr(row_ptr),
                                   _lib.ptr(col_ptr), _lib.ptr(edge_src), _lib.ptr(edge_dst), _lib.ptr(edge_w),
                                   _lib.ptr(csc_eid), _lib.ptr(uid), _lib.ptr(urows), _lib.ptr(counts), _lib.ptr(ws),
                                   ws.numel(), _lib.stream()), "prep_edges")
    el = EdgeList(B, N, ef, None, row_ptr, col_ptr, edge_src, edge_dst, edge_w, None, csc_eid)
    el.Ecap = Ecap
    el.slab = slab_t
    type_ptr, type_eid, type_pos, done = _type_sort_on_side_lane(uid, counts, Ecap, Ucap, dev, out=sort_out)
    ti = TypedInfo(uid[:Ecap], urows, counts, type_ptr, type_eid[:Ecap], type_pos[:Ecap], None, Ucap)
    ti.sort_event = done
    el._typed = ti
    _CAPTURED_COUNTS.append(counts)
    return el


def prep_edges_slab(bfm, adj):
    """Capacity-mode edge list of (bfm, adj) with all arrays in one allocation (`el.slab`); raises where the one-launch
    compaction does not serve the shape."""
    lib = _lib.load()
    if _CAPACITY is None:
        raise RuntimeError("mpnn_b200.prep_edges_slab: only inside `with graph.capacities(...)`")
    if not (bfm.is_cuda and bfm.dtype == torch.float32 and bfm.is_contiguous()
            and adj is not None and adj.dtype == torch.float32 and adj.is_contiguous()):
        raise RuntimeError("mpnn_b200.prep_edges_slab: contiguous float32 CUDA bfm / adj required")
    B, N, _, ef = bfm.shape
    if not (FUSED_PREP and lib.mpnn_prep_supported(B, N, ef, _CAPACITY[1])):
        raise RuntimeError("mpnn_b200.prep_edges_slab: batch shape not served by the one-launch compaction")
    return _prep_edges(bfm.detach(), adj.detach(), B, N, ef, _CAPACITY[0], _CAPACITY[1], slab=True)


def compact_edges(bfm, adj=None, dedup=True):
    """Compacts (bfm, adj) -> EdgeList.  One 4-byte device->host read (the edge count) sizes the arrays
    (none in capacity mode)."""
    lib = _lib.load()
    if not bfm.is_cuda:
        raise RuntimeError("mpnn_b200.compact_edges needs CUDA tensors (no CPU fallback)")
    bfm_c = _lib.f32c(bfm.detach())
    adj_c = _lib.f32c(adj.detach()) if adj is not None else None
    B, N, N2, ef = bfm_c.shape
    assert N == N2, "bfm must be [B,N,N,ef]"
    if adj_c is not None:
        assert tuple(adj_c.shape) == (B, N, N), "adj must be [B,N,N]"
    dev = bfm_c.device
    n_rows = B * N
    if _CAPACITY is not None and dedup and FUSED_PREP and lib.mpnn_prep_supported(B, N, ef, _CAPACITY[1]):
        return _prep_edges(bfm_c, adj_c, B, N, ef, _CAPACITY[0], _CAPACITY[1])
    row_ptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    col_ptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.mpnn_compact_workspace_bytes(B, N), dev)
    _lib.check(lib.mpnn_compact_count(_lib.ptr(bfm_c), _lib.ptr(adj_c), B, N, ef, _lib.ptr(row_ptr), _lib.ptr(col_ptr),
                                      _lib.ptr(ws), ws.numel(), _lib.stream()), "compact_count")
    if _CAPACITY is not None:
        Ecap = _CAPACITY[0]
        edge_src = torch.empty(Ecap, dtype=torch.int32, device=dev)
        edge_dst = torch.empty(Ecap, dtype=torch.int32, device=dev)
        csc_eid = torch.empty(Ecap, dtype=torch.int32, device=dev)
        edge_w = torch.empty(Ecap, dtype=torch.float32, device=dev)
        rows = torch.zeros(Ecap + 1, ef, dtype=torch.float32, device=dev)
        _lib.check(lib.mpnn_compact_fill(_lib.ptr(bfm_c), _lib.ptr(adj_c), B, N, ef, _lib.ptr(row_ptr),
                                         _lib.ptr(col_ptr), Ecap, _lib.ptr(edge_src), _lib.ptr(edge_dst),
                                         _lib.ptr(edge_w), _lib.ptr(rows), _lib.ptr(csc_eid), _lib.ptr(ws),
                                         _lib.stream()), "compact_fill")
        # consumers walk row_ptr / col_ptr ranges: clamp them to the allocated slots (an overflowing batch computes on a
        # truncated edge list, never out of bounds); the true count goes into the overflow flag below
        e_true = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.check(lib.mpnn_compact_clamp(_lib.ptr(row_ptr), _lib.ptr(col_ptr), n_rows, Ecap, _lib.ptr(e_true),
                                          _lib.stream()), "compact_clamp")
        el = EdgeList(B, N, ef, None, row_ptr, col_ptr, edge_src, edge_dst, edge_w, rows, csc_eid)
        el.Ecap = Ecap
        el.e_true = e_true
        if dedup:
            el._typed = dedup_rows(el, unique_capacity=_CAPACITY[1])
            _note_edge_overflow(el._typed.counts, e_true, Ecap)
            _CAPTURED_COUNTS.append(el._typed.counts)
        return el
    E = int(row_ptr[-1].item())
    STATS["E"] = max(STATS["E"], E)
    edge_src = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    edge_dst = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    csc_eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    edge_w = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    rows = torch.zeros(E + 1, ef, dtype=torch.float32, device=dev)
    _lib.check(lib.mpnn_compact_fill(_lib.ptr(bfm_c), _lib.ptr(adj_c), B, N, ef, _lib.ptr(row_ptr), _lib.ptr(col_ptr), E,
                                     _lib.ptr(edge_src), _lib.ptr(edge_dst), _lib.ptr(edge_w), _lib.ptr(rows),
                                     _lib.ptr(csc_eid), _lib.ptr(ws), _lib.stream()), "compact_fill")
    return EdgeList(B, N, ef, E, row_ptr, col_ptr, edge_src[:E], edge_dst[:E], edge_w[:E], rows, csc_eid[:E])


class TypedInfo(object):
    """uid [E] (distinct-row id per edge), urows [Ucap+1, ef] (distinct rows; row Ucap is the all-zero row x_0),
    counts [4] device {E, U, overflow, 0}, type_ptr [Ucap+1] / type_eid [E] / type_pos [E] (edges grouped by uid,
    stable).  U == Ucap unless explicit capacities were given (CUDA-graph capture)."""

    def __init__(self, uid, urows, counts, type_ptr, type_eid, type_pos, U, Ucap):
        self.uid, self.urows, self.counts = uid, urows, counts
        self.type_ptr, self.type_eid, self.type_pos = type_ptr, type_eid, type_pos
        self.U, self.Ucap = U, Ucap
        self.zero_type = Ucap
        self._tc_plan = None
        self.sort_event = None    # set when the type sort was enqueued on a side stream (capacity mode)

    def wait_sorted(self):
        """type_ptr / type_eid / type_pos are about to be read on the current stream"""
        if self.sort_event is not None:
            torch.cuda.current_stream(self.uid.device).wait_event(self.sort_event)

    def tc_plan(self, el):
        """single-type tiles of <= 128 type-sorted edges for the tcgen05 kernels (csrc/tc_message.cu), built once"""
        self.wait_sorted()
        if self._tc_plan is None:
            lib = _lib.load()
            if self.type_ptr is None:
                raise RuntimeError("mpnn_b200: the tensor-core typed path needs the type-sorted edge list")
            plan = _lib.workspace(lib.mpnn_tc_plan_bytes(el.Ecap, self.Ucap), self.uid.device)
            _lib.check(lib.mpnn_tc_plan(_lib.ptr(self.type_ptr), _lib.ptr(self.type_eid), _lib.ptr(el.edge_src),
                                        _lib.ptr(el.edge_dst), _lib.ptr(el.edge_w), el.Ecap, self.Ucap, _lib.ptr(plan),
                                        plan.numel(), _lib.stream()), "tc_plan")
            self._tc_plan = plan
        return self._tc_plan


SORT_ON_SIDE_STREAM = os.environ.get("MPNN_B200_SORT_SIDE_STREAM", "1") != "0"
TYPED_MAX_UNIQUE = int(os.environ.get("MPNN_B200_TYPED_MAX_UNIQUE", "1024"))   # beyond this many distinct bond rows the per-edge contraction (csrc/message.cu) is used


def dedup_rows(el, unique_capacity=None):
    """Exact de-duplication of el.rows by bit pattern.  Eager mode (unique_capacity None) reads U back (one 16-byte
    D2H) to size the distinct-row arrays exactly; with an explicit capacity nothing is read back (graph capture)
    and counts[2] flags an overflow."""
    lib = _lib.load()
    dev = el.rows.device
    Ecap, ef = el.Ecap, el.ef
    n_edges_ptr = el.row_ptr[el.n_rows:]
    uid = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    type_eid = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    type_pos = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    if unique_capacity is None:
        ucap0 = max(Ecap, 1)
        urows = torch.zeros(ucap0 + 1, ef, dtype=torch.float32, device=dev)
        ws = _lib.workspace(lib.mpnn_dedup_workspace_bytes(Ecap, 1), dev)
        _lib.check(lib.mpnn_dedup_rows(_lib.ptr(el.rows), _lib.ptr(n_edges_ptr), Ecap, ef, ucap0, _lib.ptr(uid),
                                       _lib.ptr(urows), _lib.ptr(counts), 0, None, None, None, _lib.ptr(ws),
                                       ws.numel(), _lib.stream()), "dedup_rows")
        U = int(counts[1].item())
        STATS["U"] = max(STATS["U"], U)
        Ucap = max(U, 1)
        urows = urows[:Ucap + 1].contiguous()   # rows >= U are zero: row Ucap is x_0
        type_ptr = None
        if Ucap <= TYPED_MAX_UNIQUE:
            type_ptr = torch.empty(Ucap + 1, dtype=torch.int32, device=dev)
            ws = _lib.workspace(lib.mpnn_type_sort_workspace_bytes(Ecap, Ucap), dev)
            _lib.check(lib.mpnn_type_sort(_lib.ptr(uid), _lib.ptr(counts), Ecap, Ucap, _lib.ptr(type_ptr),
                                          _lib.ptr(type_eid), _lib.ptr(type_pos), _lib.ptr(ws), ws.numel(),
                                          _lib.stream()), "type_sort")
        return TypedInfo(uid[:Ecap], urows, counts, type_ptr, type_eid[:Ecap], type_pos[:Ecap], U, Ucap)
    Ucap = int(unique_capacity)
    urows = torch.zeros(Ucap + 1, ef, dtype=torch.float32, device=dev)
    type_ptr = torch.empty(Ucap + 1, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.mpnn_dedup_workspace_bytes(Ecap, Ucap), dev)
    if not SORT_ON_SIDE_STREAM:
        _lib.check(lib.mpnn_dedup_rows(_lib.ptr(el.rows), _lib.ptr(n_edges_ptr), Ecap, ef, Ucap, _lib.ptr(uid),
                                       _lib.ptr(urows), _lib.ptr(counts), 1, _lib.ptr(type_ptr), _lib.ptr(type_eid),
                                       _lib.ptr(type_pos), _lib.ptr(ws), ws.numel(), _lib.stream()), "dedup_rows")
        return TypedInfo(uid[:Ecap], urows, counts, type_ptr, type_eid[:Ecap], type_pos[:Ecap], None, Ucap)
    # The forward pass needs the distinct rows and the per-edge type ids only; the edges grouped by type are read by
    # the table-gradient kernels of the backward pass (and the tensor-core plan).  The grouping (histogram, scan, fill:
    # three dependent launches) therefore runs on a side stream, a parallel branch of the captured step.
    from . import functional
    _lib.check(lib.mpnn_dedup_rows(_lib.ptr(el.rows), _lib.ptr(n_edges_ptr), Ecap, ef, Ucap, _lib.ptr(uid),
                                   _lib.ptr(urows), _lib.ptr(counts), 0, None, None, None, _lib.ptr(ws), ws.numel(),
                                   _lib.stream()), "dedup_rows")
    main = torch.cuda.current_stream(dev)
    _, side = functional._side_stream(dev, lane=5)
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        ws2 = _lib.workspace(lib.mpnn_type_sort_workspace_bytes(Ecap, Ucap), dev)
        _lib.check(lib.mpnn_type_sort(_lib.ptr(uid), _lib.ptr(counts), Ecap, Ucap, _lib.ptr(type_ptr),
                                      _lib.ptr(type_eid), _lib.ptr(type_pos), _lib.ptr(ws2), ws2.numel(),
                                      _lib.stream()), "type_sort")
        done = torch.cuda.Event()
        done.record(side)
    for t in (uid, counts, type_ptr, type_eid, type_pos):
        t.record_stream(side)
    functional._note_forward_side_work(dev, lane=5)
    ti = TypedInfo(uid[:Ecap], urows, counts, type_ptr, type_eid[:Ecap], type_pos[:Ecap], None, Ucap)
    ti.sort_event = done
    return ti


# ---- small identity cache: the same (bfm, adj) pair is compacted once per batch, whatever number of
# ---- EdgeNetworks / steps consume it (reference models call mf(afm, bfm) T times per forward).
_CACHE = collections.OrderedDict()
_CACHE_SIZE = 2     # (each entry keeps one padded batch alive: ~1 GB at BASELINE config 5)


def _key(t):
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), t.device.index)


_PINNED = {}     # (bfm address, adj address) -> [EdgeList, uses]: edge lists prepared AHEAD of the step that consumes them


def pin(bfm, adj, el):
    """`edge_list_for(bfm, adj)` returns `el` (whatever the tensors' versions, across `clear_cache()`) until `unpin()`:
    graphs.GraphedStep(pipeline_prep=True) prepares a batch's edge list one step ahead."""
    _PINNED[(bfm.data_ptr(), adj.data_ptr())] = [el, 0]


def unpin():
    uses = sum(v[1] for v in _PINNED.values())
    _PINNED.clear()
    return uses


def edge_list_for(bfm, adj=None):
    if isinstance(bfm, TypedBonds):
        return bfm.edge_list(adj)
    if _PINNED and adj is not None:
        hit = _PINNED.get((bfm.data_ptr(), adj.data_ptr()))
        if hit is not None:
            hit[1] += 1
            return hit[0]
    k = (_key(bfm), _key(adj))
    hit = _CACHE.get(k)
    if hit is not None:
        _CACHE.move_to_end(k)
        return hit[0]
    el = compact_edges(bfm, adj)
    _CACHE[k] = (el, bfm, adj)  # keep the tensors alive so data_ptr cannot be recycled under the key
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return el


def int_typed_edge_list(bt, n_types):
    """Edge list of an INTEGER bond-type tensor bt [B,N,N] (reference ggnn_msg_pass.py: 0 = no bond, t >= 1 selects
    matrix t-1): the pairs with a bond, compacted like any other batch; the type id of an edge is its bond type - 1, the
    zero type is n_types.  Cached by tensor identity like `edge_list_for`."""
    k = ("int", _key(bt), int(n_types))
    hit = _CACHE.get(k)
    if hit is not None:
        _CACHE.move_to_end(k)
        return hit[0]
    lib = _lib.load()
    x = bt.detach().to(torch.float32).unsqueeze(-1).contiguous()
    el = compact_edges(x, None, dedup=False)
    dev = x.device
    Ecap = el.Ecap
    uid = (el.rows[:max(Ecap, 1), 0].to(torch.int32) - 1).clamp_(0, n_types).contiguous()
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    counts[0:1].copy_(el.row_ptr[el.n_rows:el.n_rows + 1])
    counts[1] = n_types
    type_ptr = torch.empty(n_types + 1, dtype=torch.int32, device=dev)
    type_eid = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    type_pos = torch.empty(max(Ecap, 1), dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.mpnn_type_sort_workspace_bytes(Ecap, n_types), dev)
    _lib.check(lib.mpnn_type_sort(_lib.ptr(uid), _lib.ptr(counts), Ecap, n_types, _lib.ptr(type_ptr), _lib.ptr(type_eid),
                                  _lib.ptr(type_pos), _lib.ptr(ws), ws.numel(), _lib.stream()), "type_sort")
    el._typed = TypedInfo(uid[:Ecap], None, counts, type_ptr, type_eid[:Ecap], type_pos[:Ecap], n_types, n_types)
    _CACHE[k] = (el, bt, None)
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return el


def clear_cache():
    _CACHE.clear()
    from . import functional
    functional._REAL_ROWS.clear()


# =====================================================================================================================
# Typed bond tensor: a [B,N,N,F] bond tensor held as its DISTINCT rows  (encoders + input BNs on the
# compacted form instead of the dense O(B N^2) tensor).
#
# The datasets' raw bond rows are categorical, so a batch holds a few dozen distinct (row, adjacency value) pairs.
# Every row-wise function of the dense tensor -- the bond encoder's Linear/Tanh layers (mpnn_functions/encoders/
# bond_autoencoder.py:7-11), the adjacency-masked batch norm `bebn` (normed_encoded_basic_model.py:68) -- maps equal
# rows to equal rows, so it can be evaluated on the distinct rows alone, weighted by how often each one occurs; the
# edge network then builds its table of matrices from the resulting rows and back-propagates into them, and autograd
# carries the gradient through the row-space BN and the stock encoder modules.  Nothing of size B N^2 F is touched.
# The object follows torch's tensor-like protocol (`__torch_function__`), so the UNCHANGED reference model file can be
# handed one in place of `bfm`; anything that is not row-wise materialises the dense tensor and carries on.
# =====================================================================================================================
def _rowwise_table():
    F = torch.nn.functional
    unary = [torch.tanh, torch.relu, torch.sigmoid, F.relu, F.tanh, F.sigmoid, F.leaky_relu, F.elu, F.selu, F.celu,
             F.gelu, F.softplus, F.silu, F.hardtanh, F.relu6, F.softsign, F.tanhshrink, F.logsigmoid, F.mish,
             torch.nn.functional.hardswish, torch.nn.functional.hardsigmoid, torch.abs, torch.neg, torch.exp]
    return set(unary)


class TypedBonds(object):
    """rows [R, F] (R = Ucap + 1; row `zero_type` = Ucap stands for every pair that is not in the edge list),
    mask values a [R] (the adjacency value of the type) and occurrence counts cnt [R] (float; all on the device)."""

    _ROWWISE = None

    def __init__(self, rows, el, a, cnt, adj_key, B, N):
        self._rows, self._el, self._a, self._cnt, self._adj_key = rows, el, a, cnt, adj_key
        self._B, self._N = B, N
        self._dense = None

    # ---- tensor-like surface -------------------------------------------------------------------------------------
    @property
    def shape(self):
        return torch.Size((self._B, self._N, self._N, self._rows.shape[1]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    device = property(lambda self: self._rows.device)
    dtype = property(lambda self: self._rows.dtype)
    is_cuda = property(lambda self: self._rows.is_cuda)
    requires_grad = property(lambda self: self._rows.requires_grad)

    def float(self):
        return self

    def contiguous(self):
        return self

    def cuda(self, *a, **k):
        return self

    def __len__(self):
        return self._B

    def __repr__(self):
        return "TypedBonds(shape=%s, distinct=%d)" % (tuple(self.shape), self._rows.shape[0])

    # ---- row space -----------------------------------------------------------------------------------------------
    def with_rows(self, rows):
        """the same pairs with new row values (a row-wise function of this tensor)"""
        import copy
        el = copy.copy(self._el)
        ti = copy.copy(self._el._typed)
        ti.urows = rows
        el._typed = ti
        el.ef = rows.shape[1]
        el.rows = None
        el._per_edge = None
        el.__dict__.pop("_table_users", None)
        return TypedBonds(rows, el, self._a, self._cnt, self._adj_key, self._B, self._N)

    def edge_list(self, adj=None):
        if adj is not None and _key(adj) != self._adj_key:
            raise RuntimeError("mpnn_b200.TypedBonds: used with a different adjacency tensor than it was built from")
        return self._el

    def dense(self):
        """[B,N,N,F], differentiable in the rows (any consumer that is not row-wise ends up here)"""
        if self._dense is None:
            el = self._el
            if el.E is None:
                raise RuntimeError("mpnn_b200.TypedBonds: the dense form is not available in capacity (graph-capture) "
                                   "mode")
            ti = el._typed
            tmap = torch.full((self._B * self._N * self._N,), ti.zero_type, dtype=torch.long, device=self.device)
            if el.E:
                pair = el.edge_dst.long() * self._N + el.edge_src.long() % self._N
                tmap[pair] = ti.uid.long()
            self._dense = self._rows.index_select(0, tmap).view(self.shape)
        return self._dense

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.dense(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        F = torch.nn.functional
        if cls._ROWWISE is None:
            cls._ROWWISE = _rowwise_table()
        x = args[0] if args else None
        if isinstance(x, TypedBonds):
            rest = list(args[1:]) + list(kwargs.values())
            plain = not any(isinstance(a, TypedBonds) for a in rest)
            if func is F.linear and plain:
                return x.with_rows(func(x._rows, *args[1:], **kwargs))
            if func in cls._ROWWISE and plain and not kwargs.get("inplace", False):
                return x.with_rows(func(x._rows, *args[1:], **kwargs))
            if func in (F.softmax, F.log_softmax, torch.softmax, torch.log_softmax) and plain:
                dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
                if dim in (-1, 3) and kwargs.get("dtype") is None:
                    log = func in (F.log_softmax, torch.log_softmax)
                    return x.with_rows((torch.log_softmax if log else torch.softmax)(x._rows, -1))
            if func in (F.dropout, torch.dropout) and plain:
                training = kwargs.get("training", args[2] if len(args) > 2 else True)
                p = kwargs.get("p", args[1] if len(args) > 1 else 0.5)
                if not training or p == 0:
                    return x

        def unwrap(a):
            if isinstance(a, TypedBonds):
                return a.dense()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            return a
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in kwargs.items()})


def _tb_binop(name):
    def f(self, other):
        return getattr(self.dense(), name)(other.dense() if isinstance(other, TypedBonds) else other)
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__matmul__", "__getitem__", "__pow__", "__eq__", "__ne__", "__lt__", "__gt__", "__le__", "__ge__"):
    setattr(TypedBonds, _n, _tb_binop(_n))
TypedBonds.__neg__ = lambda self: -self.dense()
TypedBonds.__hash__ = object.__hash__


def typed_bonds(bfm, adj):
    """(bfm [B,N,N,ef] data tensor, adj [B,N,N]) -> TypedBonds, or `bfm` itself when the batch holds more than
    TYPED_MAX_UNIQUE distinct (bond row, adjacency value) pairs.  Types are keyed on the row AND the adjacency value,
    so adjacency-masked statistics are exact for weighted adjacencies too."""
    import copy
    if isinstance(bfm, TypedBonds):
        return bfm
    if bfm.requires_grad:
        raise RuntimeError("mpnn_b200.typed_bonds: bfm must be a data tensor (no gradient)")
    # (capacity mode de-duplicates eagerly; the rows are re-keyed with the adjacency value below, so skip that one)
    el0 = compact_edges(bfm, adj, dedup=False) if _CAPACITY is not None else edge_list_for(bfm, adj)
    B, N, ef = el0.B, el0.N, el0.ef
    dev = bfm.device
    Ecap = el0.Ecap
    el = copy.copy(el0)
    w = el0.edge_w if el0.E is None else el0.edge_w[:Ecap]
    aug = torch.zeros(Ecap + 1, ef + 1, dtype=torch.float32, device=dev)
    if Ecap:
        aug[:Ecap, :ef] = el0.rows[:Ecap]
        aug[:Ecap, ef] = w[:Ecap]
    el.rows, el.ef, el._typed, el._per_edge = aug, ef + 1, None, None
    el.__dict__.pop("_table_users", None)
    if _CAPACITY is not None:
        ti = dedup_rows(el, unique_capacity=_CAPACITY[1])
        if getattr(el0, "e_true", None) is not None:
            _note_edge_overflow(ti.counts, el0.e_true, Ecap)
        _CAPTURED_COUNTS.append(ti.counts)
    else:
        ti = dedup_rows(el)
    if ti.type_ptr is None:
        return bfm
    el._typed = ti
    ti.wait_sorted()                                   # the occurrence counts come from the grouping by type
    ua = ti.urows                                      # [Ucap+1, ef+1]; rows >= U are zero
    rows = ua[:, :ef].contiguous()
    a = ua[:, ef].contiguous()
    tp = ti.type_ptr
    cnt = (tp[1:] - tp[:-1]).to(torch.float32)
    n_edges = el0.row_ptr[el0.n_rows:el0.n_rows + 1].to(torch.float32)
    zero_cnt = float(B) * N * N - n_edges
    cnt = torch.cat([cnt, zero_cnt])
    tb = TypedBonds(rows, el, a, cnt, _key(adj), B, N)
    return tb.with_rows(rows)


class GatherEdgeRows(torch.autograd.Function):
    """rows[E+1, ef] as a differentiable function of the dense bfm (needed when bfm comes out of a trainable
    bond encoder, normed_encoded_basic_model.py:68): backward scatters d rows back to [B,N,N,ef]."""

    @staticmethod
    def forward(ctx, bfm, el):
        ctx.el = el
        ctx.shape = tuple(bfm.shape)
        return el.rows

    @staticmethod
    def backward(ctx, d_rows):
        el = ctx.el
        lib = _lib.load()
        dense = torch.zeros(ctx.shape, dtype=torch.float32, device=d_rows.device)
        d_rows = _lib.f32c(d_rows)
        if el.E:
            _lib.check(lib.mpnn_scatter_edge_rows(_lib.ptr(d_rows), _lib.ptr(el.edge_dst), _lib.ptr(el.edge_src), el.E,
                                                  el.N, el.ef, _lib.ptr(dense), _lib.stream()), "scatter_edge_rows")
        return dense, None
