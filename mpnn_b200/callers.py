"""Caller loops for the hot path, used by the GPU parity tests, smoke() and bench.py.

The reference's model files (models/*.py) are the real callers and run unchanged against
`mpnn_b200.mpnn_functions` (see `dropin.install`); they are not in this repository and do not exist on the
GPU box, so this single class restates the T-step composition loops behind the same constructor-injection
signature (`message_func=, message_agg_func=, update_func=, readout_func=` + `*_opts`, reference
models/basic_model.py:7-32) with the same sub-module names, i.e. the same state_dict keys:

  variant         reference loop                                   sub-modules
  "basic"         basic_model.py:50-58                             mf, ma, uf, of
  "autoencoder"   basic_graph_autoencoder.py:34-42 (encode)        mf, ma, uf, of
  "normed"        normed_basic_model.py:56-59                      mf0.., ma, uf, of, bn
  "att"           att_model.py:56-59                               mf0.., ma, uf, of, bn
  "lipo"          lipo_basic_model.py:81-86                        mf, ma (unused), uf, of, bn, ma_bn
  "normed_encoded" normed_encoded_basic_model.py:67-72             mf0.., bn0.., ma_bn0.., ma, uf, of, aebn, bebn, ae, be
"""
from torch import nn
import torch

from .modules import (AdjMsgAgg, EdgeNetwork, GraphLevelOutput, GRUUpdate, MaskBatchNorm, MaskBatchNorm1d)

from .graph import typed_bonds
import os

_PER_STEP = ("normed", "att", "normed_encoded")
TYPED_BONDS = os.environ.get("MPNN_B200_TYPED_BONDS", "1") != "0"


class MessagePassingModel(nn.Module):
    def __init__(self, variant, node_features, edge_features, message_features, adjacency_dim, output_dim,
                 message_func=EdgeNetwork, message_opts=None, message_agg_func=AdjMsgAgg, agg_opts=None,
                 update_func=GRUUpdate, update_opts=None, message_steps=3, readout_func=GraphLevelOutput,
                 readout_opts=None, atom_encoder=None, bond_encoder=None):
        super(MessagePassingModel, self).__init__()
        self.variant = variant
        message_opts = dict(message_opts or {}, node_features=node_features, edge_features=edge_features,
                            message_features=message_features)
        agg_opts = dict(agg_opts or {}, adj_dim=adjacency_dim)
        update_opts = dict(update_opts or {}, node_features=node_features, message_features=message_features)
        readout_opts = dict(readout_opts or {}, node_features=node_features, output_dim=output_dim)
        self.out_dim = output_dim
        self.iters = message_steps
        if variant in _PER_STEP:
            self.mfs = []
            for i in range(message_steps):
                self.mfs.append(message_func(**message_opts))
                self.add_module('mf' + str(i), self.mfs[-1])
        else:
            self.mf = message_func(**message_opts)
        if variant == "normed_encoded":
            self.bns, self.ma_bns = [], []
            for i in range(message_steps):
                self.bns.append(MaskBatchNorm1d(node_features))
                self.add_module('bn' + str(i), self.bns[-1])
                self.ma_bns.append(MaskBatchNorm1d(message_features))
                self.add_module('ma_bn' + str(i), self.ma_bns[-1])
            self.aebn = MaskBatchNorm1d(node_features)
            self.bebn = MaskBatchNorm1d(edge_features)
            self.ae, self.be = atom_encoder, bond_encoder
        elif variant == "lipo":
            self.bn = MaskBatchNorm1d(node_features)
            self.ma_bn = MaskBatchNorm1d(message_features)
        elif variant in ("normed", "att"):
            self.bn = MaskBatchNorm()
        self.ma = message_agg_func(**agg_opts)
        self.uf = update_func(**update_opts)
        self.of = readout_func(**readout_opts)

    def forward(self, afm, bfm, adj, mask):
        v = self.variant
        if v == "normed_encoded":
            afm = self.aebn(self.ae(afm), mask)
            if TYPED_BONDS and torch.is_tensor(bfm) and bfm.is_cuda and not bfm.requires_grad:
                # encoder + bebn + edge networks on the distinct (bond row, adjacency value) pairs (graph.TypedBonds);
                # the reference model file gets the same by being handed `typed_bonds(bfm, adj)` in place of `bfm`
                bfm = typed_bonds(bfm, adj)
            bfm = self.bebn(self.be(bfm), adj)
        node_state = afm
        if v == "basic":
            for i in range(self.iters):
                node_state = self.uf(self.ma(self.mf(afm, bfm, reuse_graph_tensors=(i > 0)), adj), node_state, mask)
        elif v == "autoencoder":
            for i in range(self.iters):
                node_state = self.uf(self.ma(self.mf(afm, bfm, reuse_graph_tensors=(i > 0)), adj), afm, mask)
        elif v in ("normed", "att"):
            for mf in self.mfs:
                node_state = self.bn(self.uf(self.ma(mf(afm, bfm), adj), node_state, mask), mask)
        elif v == "lipo":
            for i in range(self.iters):
                node_state = self.bn(self.uf(self.ma_bn(self.mf(afm, bfm, 0 != i), mask), node_state, mask), mask)
        elif v == "normed_encoded":
            for mf, bn, ma_bn in zip(self.mfs, self.bns, self.ma_bns):
                node_state = bn(self.uf(ma_bn(self.ma(mf(afm, bfm), adj), mask), node_state, mask), mask)
        else:
            raise ValueError("unknown variant %r" % (v,))
        return self.of(torch.cat([node_state, afm], dim=-1), mask=mask)


def kaiming_init(m):
    """reference init_weights (lipo_basic_model.py:88-97), nn.Linear branch."""
    if type(m) == nn.Linear:
        torch.nn.init.kaiming_uniform_(m.weight, nonlinearity='relu')
        if m.bias is not None:
            nn.init.constant_(m.bias, 0.0)
