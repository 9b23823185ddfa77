"""Device-side collate: the step right before the hot path.

The reference pads every graph to the largest one ON THE HOST (`collate_2d_graphs`, pre_process/data_loader.py:50-70)
and copies the padded tensors -- `bfm [B,N,N,ef]` is almost all zeros (757 MB at BASELINE config-5 size).  Here the
host builds a RAGGED batch (feature rows of the real atoms + the list of atom pairs that carry a bond or an adjacency
value, in the row-major (b, i, j) order the on-device compaction would find them in), ships that, and one call writes
the reference's padded layout on the GPU (`mpnn_collate_ragged`), so the modules keep their reference signatures.
"""
import numpy as np
import torch

from . import _lib


class RaggedBatch(object):
    """Host-side (pinned) ragged form of a list of graphs {afm [n,Fa], bfm [n,n,ef], adj [n,n]}."""

    def __init__(self, B, N, Fa, ef, atom_row, afm_cat, edge_dst, edge_j, edge_w, edge_x, labels=None):
        self.B, self.N, self.Fa, self.ef = B, N, Fa, ef
        self.atom_row, self.afm_cat = atom_row, afm_cat
        self.edge_dst, self.edge_j, self.edge_w, self.edge_x = edge_dst, edge_j, edge_w, edge_x
        self.labels = labels

    @staticmethod
    def from_graphs(graphs, labels=None, pin=True):
        B = len(graphs)
        N = max(g["afm"].shape[0] for g in graphs)
        Fa = graphs[0]["afm"].shape[1]
        ef = graphs[0]["bfm"].shape[2]
        rows, feats, dst, jj, ww, xx = [], [], [], [], [], []
        for b, g in enumerate(graphs):
            n = g["afm"].shape[0]
            rows.append(b * N + np.arange(n, dtype=np.int32))
            feats.append(np.asarray(g["afm"], np.float32))
            keep = (np.asarray(g["bfm"]) != 0).any(-1) | (np.asarray(g["adj"]) != 0)
            i, j = np.nonzero(keep)                      # row-major: the order torch.nonzero / the compaction use
            dst.append((b * N + i).astype(np.int32))
            jj.append(j.astype(np.int32))
            ww.append(np.asarray(g["adj"], np.float32)[i, j])
            xx.append(np.asarray(g["bfm"], np.float32)[i, j])

        def cat(parts, shape, dtype):
            a = np.concatenate(parts, 0) if parts else np.zeros(shape, dtype)
            t = torch.from_numpy(np.ascontiguousarray(a.astype(dtype, copy=False)))
            return t.pin_memory() if pin and torch.cuda.is_available() else t

        lab = None
        if labels is not None:
            lab = torch.from_numpy(np.ascontiguousarray(np.asarray(labels, np.float32)))
            lab = lab.pin_memory() if pin and torch.cuda.is_available() else lab
        return RaggedBatch(B, N, Fa, ef, cat(rows, (0,), np.int32), cat(feats, (0, Fa), np.float32),
                           cat(dst, (0,), np.int32), cat(jj, (0,), np.int32), cat(ww, (0,), np.float32),
                           cat(xx, (0, ef), np.float32), lab)

    def tensors(self):
        t = dict(atom_row=self.atom_row, afm_cat=self.afm_cat, edge_dst=self.edge_dst, edge_j=self.edge_j,
                 edge_w=self.edge_w, edge_x=self.edge_x)
        if self.labels is not None:
            t["labels"] = self.labels
        return t

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.tensors().values())

    def to(self, device, non_blocking=True):
        """copies the ragged arrays to `device` (no padding yet)"""
        d = {k: v.to(device, non_blocking=non_blocking) for k, v in self.tensors().items()}
        return RaggedBatch(self.B, self.N, self.Fa, self.ef, d["atom_row"], d["afm_cat"], d["edge_dst"], d["edge_j"],
                           d["edge_w"], d["edge_x"], d.get("labels"))

    def scatter_padded(self, out=None):
        """On a CUDA ragged batch: writes the reference's padded tensors (into `out` = dict of preallocated afm/bfm/adj/
        mask if given) and returns the batch dict the reference's models consume."""
        lib = _lib.load()
        if not self.afm_cat.is_cuda:
            raise RuntimeError("mpnn_b200.loader: scatter_padded needs the ragged batch on a CUDA device (use .to())")
        dev = self.afm_cat.device
        B, N, Fa, ef = self.B, self.N, self.Fa, self.ef
        if out is None:
            out = dict(afm=torch.empty(B, N, Fa, dtype=torch.float32, device=dev),
                       bfm=torch.empty(B, N, N, ef, dtype=torch.float32, device=dev),
                       adj=torch.empty(B, N, N, dtype=torch.float32, device=dev),
                       mask=torch.empty(B, N, 1, dtype=torch.float32, device=dev))
        for k, shp in (("afm", (B, N, Fa)), ("bfm", (B, N, N, ef)), ("adj", (B, N, N)), ("mask", (B, N, 1))):
            if tuple(out[k].shape) != shp or not out[k].is_contiguous():
                raise RuntimeError("mpnn_b200.loader: output %s must be a contiguous %s tensor" % (k, shp))
        _lib.check(lib.mpnn_collate_ragged(_lib.ptr(self.atom_row), _lib.ptr(self.afm_cat), self.afm_cat.shape[0], Fa,
                                           _lib.ptr(self.edge_dst), _lib.ptr(self.edge_j), _lib.ptr(self.edge_w),
                                           _lib.ptr(self.edge_x), self.edge_dst.shape[0], ef, B, N, _lib.ptr(out["afm"]),
                                           _lib.ptr(out["bfm"]), _lib.ptr(out["adj"]), _lib.ptr(out["mask"]),
                                           _lib.stream()), "collate_ragged")
        if self.labels is not None and "labels" not in out:
            out["labels"] = self.labels
        return out


def collate_ragged(graphs, labels=None, device=None):
    """Drop-in for the reference's `collate_2d_graphs` + `from_numpy(...).cuda()` (data_loader.py:50-70, utils.py:5-13):
    same padded batch dict, built on the GPU from a ragged host->device transfer."""
    rb = RaggedBatch.from_graphs(graphs, labels)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return rb.to(device).scatter_padded()
