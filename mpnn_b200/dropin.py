"""Makes the reference's model files import THIS implementation.

The reference models do `from mpnn_functions import *`, `from mask_batch_norm import MaskBatchNorm[1d]` and
`from mpnn_functions.message.ggnn_msg_pass import GGNNMsgPass` (models/basic_model.py:3,
normed_basic_model.py:4, lipo_basic_model.py:3-5).  After `install()` those names resolve to the CUDA-backed
modules, so the model files run unchanged.  (Constructor injection -- message_func=..., update_func=... --
needs no install at all.)
"""
import importlib
import sys

_ALIASES = {
    "mpnn_functions": "mpnn_b200.mpnn_functions",
    "mpnn_functions.message": "mpnn_b200.mpnn_functions.message",
    "mpnn_functions.message.ggnn_msg_pass": "mpnn_b200.mpnn_functions.message.ggnn_msg_pass",
    "mpnn_functions.message.edge_network": "mpnn_b200.mpnn_functions.message.edge_network",
    "mpnn_functions.message.att_edge_network": "mpnn_b200.mpnn_functions.message.att_edge_network",
    "mpnn_functions.message_aggregators": "mpnn_b200.mpnn_functions.message_aggregators",
    "mpnn_functions.update": "mpnn_b200.mpnn_functions.update",
    "mpnn_functions.readout": "mpnn_b200.mpnn_functions.readout",
    "mask_batch_norm": "mpnn_b200.mask_batch_norm",
}


def install():
    for alias, real in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(real)


def uninstall():
    for alias in _ALIASES:
        sys.modules.pop(alias, None)
