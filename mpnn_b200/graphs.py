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
    """

    def __init__(self, step_fn, example_inputs, warmup=3, headroom=1.25, edge_capacity=None, unique_capacity=None):
        self.static = {k: v.clone() for k, v in example_inputs.items()}
        self.step_fn = step_fn
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
        with graph.capacities(self.edge_capacity, self.unique_capacity):
            with torch.cuda.graph(self.graph):
                # every zero-initialised buffer of the step is a slice of this arena: one memset node, no fill launches
                _lib.check(lib.mpnn_zero_bytes(_lib.ptr(self.arena), self.arena.numel(), _lib.stream()), "zero_bytes")
                functional.arena_begin(self.arena)
                try:
                    self.loss = step_fn(self.static)
                finally:
                    functional.arena_end()
                functional.join_side_streams()   # no forked branch may outlive the capture
        self._counts = list(graph._CAPTURED_COUNTS)
        del graph._CAPTURED_COUNTS[:]
        graph.clear_cache()

    def load(self, inputs, non_blocking=True):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.loss

    def __call__(self, inputs=None):
        if inputs is not None:
            self.load(inputs)
        return self.replay()

    def check(self):
        """One small D2H read: did any batch replayed since the last check exceed the captured edge / distinct-row
        capacities?  (Kernels clamp their reads to the capacities, so an overflowing batch computes on a truncated edge
        list; the step's result must be discarded by the caller.)"""
        for c in self._counts:
            e, u, flag, _ = c.cpu().tolist()
            if flag:
                c[2] = 0     # the flag is sticky on the device (set by any replay since the last check): clear it here
                raise RuntimeError("mpnn_b200.GraphedStep: batch with %d edges / %d distinct bond rows exceeds the "
                                   "captured capacities (%d / %d); re-capture with larger capacities"
                                   % (e, u, self.edge_capacity, self.unique_capacity))
