"""Makes the reference's model files import THIS implementation.

The reference models do `from mpnn_functions import *`, `from mask_batch_norm import MaskBatchNorm[1d]`,
`from mpnn_functions.message.ggnn_msg_pass import GGNNMsgPass` (models/basic_model.py:3,
normed_basic_model.py:4, lipo_basic_model.py:3-5) and the drivers do
`from mpnn_functions.encoders.bond_autoencoder import BondAutoEncoder` (test_graph_encode_norm_ecfp.py:20-21).
After `install()` those names resolve to the CUDA-backed modules, so the model files run unchanged.
(Constructor injection -- message_func=..., update_func=... -- needs no install at all.)

`reference_models(dir)` imports the UNMODIFIED model files of a reference checkout (or of the byte-identical test
fixtures under tests/ref_models/) against the installed aliases; `reference_model(variant, ...)` builds one of them.
The model files are clients of this package's boundary: nothing in mpnn_b200 depends on them.
"""
import importlib
import importlib.util
import os
import sys
import types

_ALIASES = {
    "mpnn_functions": "mpnn_b200.mpnn_functions",
    "mpnn_functions.message": "mpnn_b200.mpnn_functions.message",
    "mpnn_functions.message.ggnn_msg_pass": "mpnn_b200.mpnn_functions.message.ggnn_msg_pass",
    "mpnn_functions.message.edge_network": "mpnn_b200.mpnn_functions.message.edge_network",
    "mpnn_functions.message.att_edge_network": "mpnn_b200.mpnn_functions.message.att_edge_network",
    "mpnn_functions.message_aggregators": "mpnn_b200.mpnn_functions.message_aggregators",
    "mpnn_functions.update": "mpnn_b200.mpnn_functions.update",
    "mpnn_functions.readout": "mpnn_b200.mpnn_functions.readout",
    "mpnn_functions.encoders": "mpnn_b200.mpnn_functions.encoders",
    "mpnn_functions.encoders.auto_encoder": "mpnn_b200.mpnn_functions.encoders.auto_encoder",
    "mpnn_functions.encoders.atom_autoencoder": "mpnn_b200.mpnn_functions.encoders.atom_autoencoder",
    "mpnn_functions.encoders.bond_autoencoder": "mpnn_b200.mpnn_functions.encoders.bond_autoencoder",
    "mask_batch_norm": "mpnn_b200.mask_batch_norm",
}

# model file -> (class, forward method) of the T-step loops this package serves
MODEL_FILES = {
    "basic": ("basic_model", "BasicModel", "forward"),
    "autoencoder": ("basic_graph_autoencoder", "Encoder", "encode"),
    "normed": ("normed_basic_model", "BasicModel", "forward"),
    "att": ("att_model", "BasicModel", "forward"),
    "lipo": ("lipo_basic_model", "BasicModel", "forward"),
    "normed_encoded": ("normed_encoded_basic_model", "BasicModel", "forward"),
    "normed_encoded_ecfp": ("normed_encoded_basic_model_ecfp", "BasicModel", "forward"),
}
_WRAPPERS = ("graph_model_wrapper", "graph_norm_wrapper", "batch_norm_graph_wrapper")

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOADED = {}


def install():
    for alias, real in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(real)


def uninstall():
    for alias in _ALIASES:
        sys.modules.pop(alias, None)


def default_models_dir():
    """$MPNN_B200_REF_MODELS, else the byte-identical fixtures shipped with the tests (tests/ref_models/models)."""
    d = os.environ.get("MPNN_B200_REF_MODELS")
    if d:
        return d
    return os.path.join(_ROOT, "tests", "ref_models", "models")


def reference_models(models_dir=None):
    """Namespace of the reference's model modules (`ns.normed_basic_model.BasicModel`, `ns.graph_model_wrapper...`),
    imported from `models_dir` UNCHANGED with this package installed as `mpnn_functions` / `mask_batch_norm`."""
    models_dir = os.path.abspath(models_dir or default_models_dir())
    if models_dir in _LOADED:
        return _LOADED[models_dir]
    if not os.path.isdir(models_dir):
        raise RuntimeError("mpnn_b200.dropin: no reference model files under %s" % models_dir)
    install()
    ns = types.SimpleNamespace()
    names = list(_WRAPPERS) + sorted(set(v[0] for v in MODEL_FILES.values()))
    saved = {}
    try:
        for name in names:
            path = os.path.join(models_dir, name + ".py")
            if not os.path.exists(path):
                continue
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            # att_model.py:4 does `from batch_norm_graph_wrapper import MaskBatchNorm` (py2 implicit-relative import):
            # the wrappers are importable as top-level names while the model files load
            saved[name] = sys.modules.get(name)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            setattr(ns, name, mod)
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    _LOADED[models_dir] = ns
    return ns


def reference_model(variant, node_features, edge_features, message_features, adjacency_dim, output_dim,
                    models_dir=None, **kw):
    """Instantiates the reference's own model class for `variant` (see MODEL_FILES) with fresh `*_opts` dicts (the
    reference's constructors mutate their shared default dicts, models/basic_model.py:8-24).  For the auto-encoder the
    instance's `forward` is its `encode` (basic_graph_autoencoder.py:34-42; its own `forward` returns None)."""
    fname, cls, method = MODEL_FILES[variant]
    mod = getattr(reference_models(models_dir), fname)
    for k in ("message_opts", "agg_opts", "update_opts", "readout_opts"):
        kw[k] = dict(kw.get(k) or {})
    model = getattr(mod, cls)(node_features, edge_features, message_features, adjacency_dim, output_dim, **kw)
    if method != "forward":
        model.forward = getattr(model, method)
    return model


def kaiming_init(m):
    """the reference's init_weights, nn.Linear branch (lipo_basic_model.py:88-97) -- for models without their own"""
    import torch
    from torch import nn
    if type(m) == nn.Linear:
        torch.nn.init.kaiming_uniform_(m.weight, nonlinearity='relu')
        if m.bias is not None:
            nn.init.constant_(m.bias, 0.0)
