"""CUDA-graph capture of a whole train (or inference) step of the message-passing path.

At the reference's batch sizes (16..256 molecules, test_lipo.py:150) every kernel of the path runs for a few
microseconds, so a step is bound by launch latency and Python dispatch, not by HBM or the tensor pipe.  The
B200 answer is to capture the step once -- compaction, de-duplication, edge networks, message passing, readout,
loss, backward, optimizer -- and replay it: no Python, no host<->device synchronisation, one graph launch.

Requirements the library meets for this: every kernel takes its data-dependent sizes (edge count E, distinct
bond rows U) from DEVICE memory, arrays are sized by capacities fixed at capture time
(`mpnn_b200.graph.capacities`), and an overflow only raises a flag (`GraphedStep.check()`).
"""
import torch

from . import _lib, functional, graph


class GraphedStep(object):
    """Captures `step_fn(inputs) -> loss` (which does zero_grad / backward / optimizer.step itself) into a CUDA
    graph over static copies of `example_inputs` (dict of CUDA tensors; all later batches must have these shapes).

        gs = GraphedStep(step_fn, batch)        # 3 eager warm-up steps on a side stream + 1 eager dry run with the
                                                # captured array sizes (it sizes the zero-buffer arena), then capture:
                                                # step_fn has run `gs.eager_steps` times when the constructor returns
        loss = gs(batch)                        # copies the batch into the static buffers, replays
        gs.check()                              # raises if a batch overflowed the captured capacities

    pipeline_prep=True software-pipelines the input preprocessing: compaction + de-duplication + type sort of a batch
    (csrc/prep.cu; they depend on `bfm` / `adj` only, not on the weights) run one replay AHEAD, as a parallel branch of the
    previous batch's train step (forked behind the fused step kernel's forward: the one-CTA-per-SM step kernels and the
    cooperative compaction kernel cannot share an SM, the readout / head kernels in between leave room), instead of at
    the head of the step's critical path.  Every replay still does one preprocessing and one train step.  In this mode
    the static `bfm` / `adj` buffers hold the batch AFTER the one being trained on; the step may use them only through
    the message-passing modules (which receive the edge list prepared one replay earlier), not read them densely:

        gs = GraphedStep(step_fn, batch0, pipeline_prep=True)    # batch0's edge list is prepared eagerly
        gs.load_next(batch1)                    # bfm / adj of the batch AFTER the one about to be trained on
        loss0 = gs.replay()                     # trains on batch0, prepares batch1's edge list
        gs.load(batch1); gs.load_next(batch2)   # batch1's afm / mask / labels ...; batch2's bfm / adj
        loss1 = gs.replay()                     # ...

    host_io=True puts the input transfers INTO the graph: `gs.host[group][1]` are pinned host tensors (views of one flat
    pinned allocation per input group) the data loader fills; every replay first moves the staged copy of the inputs
    into the static buffers (device copy), then runs the step while a parallel branch of the same graph copies the host
    buffers into the staging area for the next replay (H2D from pinned memory, overlapped with the step's kernels).
    A train step is then ONE graph launch for the host, transfers included.  The host buffers must hold, when replay k
    is launched, what replay k + 1 consumes (with pipeline_prep: bfm / adj of batch k + 2, the rest of batch k + 1).
    """

    def __init__(self, step_fn, example_inputs, warmup=3, headroom=1.25, edge_capacity=None, unique_capacity=None,
                 pipeline_prep=False, host_io=False):
        # static inputs as views of two flat allocations ("bonds": bfm / adj, "rest": everything else), so a batch can be
        # loaded with two device copies whatever its number of tensors (`flat`, `mirror()`)
        dev0 = next(iter(example_inputs.values())).device
        self.layout, self.flat, self.static = {}, {}, {}
        for group, ks in (("bonds", [k for k in example_inputs if k in ("bfm", "adj")]),
                          ("rest", [k for k in example_inputs if k not in ("bfm", "adj")])):
            self.layout[group], off = [], 0
            for k in ks:
                v = example_inputs[k]
                self.layout[group].append((k, off, tuple(v.shape), v.dtype))
                off += (v.numel() * v.element_size() + 255) // 256 * 256
            self.layout[group] = (self.layout[group], off)
            self.flat[group], views = self.mirror(group, device=dev0)
            for k in ks:
                views[k].copy_(example_inputs[k])
            self.static.update(views)
        self.step_fn = step_fn
        self.pipeline_prep = bool(pipeline_prep)
        self.host_io = bool(host_io)
        self.host, self.stage = {}, {}
        if self.host_io:
            for group in ("bonds", "rest"):
                self.host[group] = self.mirror(group, device="cpu", pin_memory=True)
                self.stage[group] = self.mirror(group, device=dev0)
                self.host[group][0].copy_(self.flat[group])
                self.stage[group][0].copy_(self.flat[group])
        self.eager_steps = warmup + 1    # executions of step_fn before the capture (warm-up + the arena dry run)
        graph.STATS["E"] = graph.STATS["U"] = graph.STATS["n_real"] = 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                graph.clear_cache()
                step_fn(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n_rows = self.static["bfm"].shape[0] * self.static["bfm"].shape[1] if "bfm" in self.static else 1
        self.edge_capacity = int(edge_capacity or max(int(graph.STATS["E"] * headroom) + 64, 64))
        self.unique_capacity = int(unique_capacity or min(max(int(graph.STATS["U"] * 1.5) + 8, 16),
                                                           graph.TYPED_MAX_UNIQUE))
        del n_rows
        # dry run with the captured array sizes: how many bytes of zero-initialised buffers does one step ask for?
        graph.clear_cache()
        functional.arena_measure()
        try:
            with graph.capacities(self.edge_capacity, self.unique_capacity):
                with torch.cuda.stream(side):
                    step_fn(self.static)
        finally:
            need = functional.arena_end()
        torch.cuda.synchronize()
        del graph._CAPTURED_COUNTS[:]
        self.arena = torch.empty(max(need, 256), dtype=torch.uint8, device=next(iter(self.static.values())).device)
        self.graph = torch.cuda.CUDAGraph()
        graph.clear_cache()
        functional._FWD_SIDE.clear()
        lib = _lib.load()
        dev = next(iter(self.static.values())).device
        # overflow flags of every replay since the last check(), OR-ed on the device (the per-step counts are cleared or
        # overwritten by the next replay)
        self._sticky = torch.zeros(1, dtype=torch.int32, device=dev)
        with graph.capacities(self.edge_capacity, self.unique_capacity):
            if self.pipeline_prep:
                # the edge list the FIRST replay trains on, built eagerly; every replay hands the next one over
                self.next = {k: self.static[k] for k in ("bfm", "adj")}
                self._el_cur = graph.prep_edges_slab(self.static["bfm"], self.static["adj"])
                torch.cuda.synchronize()
                del graph._CAPTURED_COUNTS[:]
                self._el_cur._typed.sort_event = None       # (recorded outside the capture; the work is finished)
            # captured on the stream the warm-up steps ran on: the per-stream persistent workspaces (_lib.clean_workspace) of
            # the warm-up are the capture's, so no zero-fill of a fresh workspace lands in the graph
            with torch.cuda.graph(self.graph, stream=side):
                # every zero-initialised buffer of the step is a slice of this arena: one memset node, no fill launches
                _lib.check(lib.mpnn_zero_bytes(_lib.ptr(self.arena), self.arena.numel(), _lib.stream()), "zero_bytes")
                functional.arena_begin(self.arena)
                try:
                    if self.host_io:
                        self._capture_host_io(dev)
                    if self.pipeline_prep:
                        self._el_nxt = None
                        graph.pin(self.static["bfm"], self.static["adj"], self._el_cur)
                        graph._CAPTURED_COUNTS.append(self._el_cur._typed.counts)   # the optimizer's overflow guard
                        functional.AFTER_CHAIN_FWD.append(lambda: self._capture_prep_branch(dev))
                    self.loss = step_fn(self.static)
                    if self.pipeline_prep and self._el_nxt is None:
                        del functional.AFTER_CHAIN_FWD[:]
                        self._capture_prep_branch(dev)      # (no fused step kernel in this step: no overlap, same result)
                finally:
                    functional.arena_end()
                    uses = graph.unpin()
                    del functional.AFTER_CHAIN_FWD[:]
                functional.join_side_streams()   # no forked branch may outlive the capture
                if not self.pipeline_prep:
                    for c in graph._CAPTURED_COUNTS:
                        self._sticky.bitwise_or_(c[2:3])
                if self.pipeline_prep:
                    if not uses:
                        raise RuntimeError("mpnn_b200.GraphedStep(pipeline_prep=True): the step never asked for the "
                                           "edge list of (inputs['bfm'], inputs['adj'])")
                    # every consumer of the current edge list has finished, the next one is complete: hand it over
                    self._el_cur.slab.copy_(self._el_nxt.slab)
        self._counts = list(graph._CAPTURED_COUNTS)
        del graph._CAPTURED_COUNTS[:]
        graph.clear_cache()

    def mirror(self, group, device=None, pin_memory=False):
        """(flat uint8 buffer, {name: tensor view}) with the layout of the static input group "bonds" (bfm, adj) or "rest":
        a host (pinned) or device staging copy of a batch that moves with ONE copy into `self.flat[group]`"""
        entries, total = self.layout[group]
        flat = torch.empty(max(total, 256), dtype=torch.uint8, device=device, pin_memory=pin_memory)
        views = {}
        for k, off, shape, dtype in entries:
            n = 1
            for d_ in shape:
                n *= d_
            nb = n * torch.empty(0, dtype=dtype).element_size()
            views[k] = flat[off:off + nb].view(dtype).view(shape)
        return flat, views

    def _capture_host_io(self, dev):
        """(inside the capture, at the head of the step) staged inputs -> static inputs, then, as a parallel branch, the
        host's pinned buffers -> staging area for the next replay.  With pipeline_prep only the preprocessing branch reads
        bfm / adj, so their (large) device copy runs on that branch and the step's own chain waits for the small one."""
        main = torch.cuda.current_stream(dev)
        root = torch.cuda.Event()
        root.record(main)
        self.flat["rest"].copy_(self.stage["rest"][0], non_blocking=True)
        bonds_done = None
        if self.pipeline_prep:
            key7, lane7 = functional._side_stream(dev, lane=7)
            lane7.wait_event(root)
            with torch.cuda.stream(lane7):
                self.flat["bonds"].copy_(self.stage["bonds"][0], non_blocking=True)
                bonds_done = torch.cuda.Event()
                bonds_done.record(lane7)
            functional._FWD_SIDE.add(key7)
        else:
            self.flat["bonds"].copy_(self.stage["bonds"][0], non_blocking=True)
        fork = torch.cuda.Event()
        fork.record(main)
        key, lane = functional._side_stream(dev, lane=8)
        lane.wait_event(fork)
        if bonds_done is not None:
            lane.wait_event(bonds_done)
        with torch.cuda.stream(lane):
            for group in ("bonds", "rest"):
                self.stage[group][0].copy_(self.host[group][0], non_blocking=True)
        functional._FWD_SIDE.add(key)

    def _capture_prep_branch(self, dev):
        """(inside the capture) the NEXT batch's compaction / de-duplication / type sort on a side lane forked from the
        current position of the capturing stream"""
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        key, lane = functional._side_stream(dev, lane=7)
        lane.wait_event(fork)
        with torch.cuda.stream(lane):
            self._sticky.bitwise_or_(self._el_cur._typed.counts[2:3])    # overflow flag of the batch being trained on
            self._el_nxt = graph.prep_edges_slab(self.next["bfm"], self.next["adj"])
        functional._FWD_SIDE.add(key)
        # the optimizer's overflow guard is the CURRENT batch's flag, not the next one's
        nxt = self._el_nxt._typed.counts
        graph._CAPTURED_COUNTS[:] = [c for c in graph._CAPTURED_COUNTS if c is not nxt]

    def load(self, inputs, non_blocking=True):
        """copies a batch into the static buffers (pipeline_prep: everything but bfm / adj, see load_next)"""
        for k, v in inputs.items():
            if self.pipeline_prep and k in ("bfm", "adj"):
                continue
            self.static[k].copy_(v, non_blocking=non_blocking)

    def load_next(self, inputs, non_blocking=True):
        """pipeline_prep: bfm / adj of the batch that FOLLOWS the one the next replay trains on"""
        for k in ("bfm", "adj"):
            self.static[k].copy_(inputs[k], non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.loss

    def __call__(self, inputs=None):
        if inputs is not None:
            if self.pipeline_prep:
                raise RuntimeError("mpnn_b200.GraphedStep(pipeline_prep=True): use load() / load_next() + replay()")
            self.load(inputs)
        return self.replay()

    def check(self):
        """One small D2H read: did any batch replayed since the last check exceed the captured edge / distinct-row
        capacities?  (Kernels clamp their reads to the capacities, so an overflowing batch computes on a truncated edge
        list; the step's result must be discarded by the caller.)"""
        if int(self._sticky.item()):
            self._sticky.zero_()
            last = [c.cpu().tolist() for c in self._counts]
            raise RuntimeError("mpnn_b200.GraphedStep: a batch replayed since the last check exceeded the captured "
                               "capacities (%d edges / %d distinct bond rows; the most recent edge lists hold %s); "
                               "re-capture with larger capacities"
                               % (self.edge_capacity, self.unique_capacity,
                                  ", ".join("%d / %d" % (c[0], c[1]) for c in last) or "-"))
