"""ctypes binding of the C-ABI library (include/mpnn_b200.h).  No torch types cross this boundary:
raw device pointers, sizes and a cudaStream_t.  There is NO fallback: if the library is missing the
import of any op raises, and every op refuses non-CUDA tensors."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmpnn_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_Z = ctypes.c_size_t
_F = ctypes.c_float
_PP = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must match include/mpnn_b200.h (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "mpnn_version": (_I, []),
    "mpnn_zero_bytes": (_I, [_P, _Z, _P]),
    "mpnn_last_error": (ctypes.c_char_p, []),
    "mpnn_set_tensor_cores": (_I, [_I]),
    "mpnn_tensor_cores_enabled": (_I, []),
    "mpnn_segment_sum": (_I, [_P, _P, _P, _I, _I, _L, _P, _L, _I, _F, _P]),
    "mpnn_colsum_workspace_bytes": (_Z, [_L, _I]),
    "mpnn_colsum": (_I, [_P, _P, _L, _I, _L, _L, _P, _I, _P, _Z, _P]),
    "mpnn_gemm_workspace_bytes": (_Z, [_I, _I, _I]),
    "mpnn_gemm": (_I, [_P, _P, _P, _I, _I, _I, _L, _L, _L, _L, _L, _P, _I, _P, _Z, _P]),
    "mpnn_compact_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_compact_count": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "mpnn_compact_fill": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "mpnn_collate_ragged": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mpnn_dedup_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_dedup_rows": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "mpnn_type_sort_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_type_sort": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "mpnn_typed_dp": (_I, [_I, _I]),
    "mpnn_graph_sum": (_I, [_P, _I, _I, _I, _P, _P]),
    "mpnn_table_from_flat": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "mpnn_table_to_flat": (_I, [_P, _I, _I, _I, _P, _P]),
    "mpnn_enet_supported": (_I, [_I, _I, _I]),
    "mpnn_enet_max_dp": (_I, []),
    "mpnn_enet_saved_floats": (_L, [_I, _I, _I]),
    "mpnn_enet_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "mpnn_enet_fwd": (_I, [_P, _I, _I, _I, _PP, _PP, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P]),
    "mpnn_enet_bwd": (_I, [_P, _I, _I, _I, _PP, _P, _I, _I, _P, _I, _I, _P, _P, _PP, _PP, _P, _P, _P, _P, _P, _Z,
                           _P]),
    "mpnn_enet_max_nets": (_I, []),
    "mpnn_enet_fwd_multi": (_I, [_I, _P, _I, _I, _I, _PP, _PP, _PP, _I, _I, _PP, _PP, _I, _I, _PP, _PP, _PP, _P]),
    "mpnn_enet_bwd_multi": (_I, [_I, _P, _I, _I, _I, _PP, _PP, _I, _I, _PP, _I, _I, _PP, _PP, _PP, _PP, _PP, _PP, _PP,
                                 _PP, _P, _Z, _P]),
    "mpnn_tmsg_bwd_table_multi": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "mpnn_tmsg_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "mpnn_tmsg_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mpnn_tmsg_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P,
                           _P, _P, _P, _Z, _P]),
    "mpnn_tc_dp": (_I, [_I, _I]),
    "mpnn_tc_plan_bytes": (_Z, [_I, _I]),
    "mpnn_tc_plan": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "mpnn_tc_edge_gemm_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_tc_edge_gemm": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _Z, _P]),
    "mpnn_tc_table_grad_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_tc_table_grad": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _I, _P, _P, _Z, _P]),
    "mpnn_tc_dense_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_tc_dense_gemm": (_I, [_P, _L, _I, _I, _I, _I, _P, _L, _L, _L, _L, _I, _I, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "mpnn_tc_gru_supported": (_I, [_I]),
    "mpnn_tc_gru_workspace_bytes": (_Z, [_I]),
    "mpnn_tc_gru_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _Z, _P]),
    "mpnn_tc_linear_supported": (_I, [_I, _I]),
    "mpnn_tc_linear_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_tc_linear_fwd": (_I, [_P, _L, _I, _I, _P, _I, _P, _P, _I, _I, _P, _Z, _P]),
    "mpnn_tc_linear_bwd_data": (_I, [_P, _L, _I, _I, _P, _I, _P, _I, _I, _P, _Z, _P]),
    "mpnn_tc_linear_bwd_weight": (_I, [_P, _L, _I, _I, _P, _I, _I, _P, _P, _Z, _P]),
    "mpnn_tc_dense_gemm_ll": (_I, [_P, _L, _I, _I, _I, _I, _P, _L, _L, _L, _L, _I, _I, _P, _P, _I, _L, _I, _I, _I, _P, _Z, _P]),
    "mpnn_tc_gru_param_workspace_bytes": (_Z, []),
    "mpnn_tc_gru_param_bias_parts": (_I, []),
    "mpnn_tc_debug": (None, [_P]),
    "mpnn_tc_gru_data_workspace_bytes": (_Z, []),
    "mpnn_tc_gru_data_grad": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _Z, _P]),
    "mpnn_tc_gru_param_point": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_gru_bwd_one_pass": (_I, [_I]),
    "mpnn_tc_gru_param_grad": (_I, [_P, _P, _P, _I, _L, _I, _P, _P, _P, _Z, _P]),
    "mpnn_tc_dense_grad_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_tc_dense_gemm_tn": (_I, [_P, _L, _I, _I, _P, _I, _I, _I, _I, _I, _P, _L, _L, _P, _Z, _P]),
    "mpnn_scatter_edge_rows": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "mpnn_edge_trunk_saved_floats": (_L, [_I, _I, _I, _I, _I, ctypes.POINTER(_L), ctypes.POINTER(_I)]),
    "mpnn_edge_trunk_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "mpnn_edge_trunk_fwd": (_I, [_P, _I, _I, _I, _PP, _PP, _P, _I, _I, _P, _P, _Z, _P]),
    "mpnn_edge_trunk_bwd": (_I, [_P, _I, _I, _I, _PP, _P, _I, _I, _P, _P, _I, _PP, _PP, _P, _P, _P, _Z, _P]),
    "mpnn_message_wt_floats": (_L, [_I, _I, _I]),
    "mpnn_message_prepare": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "mpnn_message_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mpnn_message_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "mpnn_message_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I,
                              _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_gru_workspace_bytes": (_Z, [_L, _I]),
    "mpnn_gru_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _Z, _P]),
    "mpnn_gru_agg_supported": (_I, [_I]),
    "mpnn_gru_fwd_agg": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "mpnn_tc_gru_fwd_agg": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "mpnn_gru_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_gru_bwd_partial_bytes": (_Z, [_L, _I]),
    "mpnn_gru_bwd_data": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "mpnn_gru_bwd_params": (_I, [_P, _I, _L, _I, _P, _P, _P, _P, _P]),
    "mpnn_bn_workspace_bytes": (_Z, [_L, _I]),
    "mpnn_mask_bn_fwd": (_I, [_P, _P, _L, _I, _F, _P, _P, _P, _Z, _P]),
    "mpnn_mask_bn_bwd": (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _Z, _P]),
    "mpnn_mask_bn1d_fwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _F, _P, _P, _P, _Z, _P]),
    "mpnn_mask_bn1d_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _P, _P, _P, _P, _Z, _P]),
    "mpnn_glo_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "mpnn_glo_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_row_bn_fwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P]),
    "mpnn_row_bn_bwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "mpnn_glo_bwd_split_supported": (_I, [_I, _I, _I]),
    "mpnn_glo_bwd_data": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "mpnn_glo_bwd_params": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mpnn_glo_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_adam_step": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _PP, _I, _P]),
    "mpnn_adam_step_ddp": (_I, [_I, _PP, _PP, _PP, _PP, _P, _P, _P, _P, _F, _F, _F, _F, _F, _PP, _PP, _L, _I, _I, _P]),
    "mpnn_compact_clamp": (_I, [_P, _P, _I, _I, _P, _P]),
    "mpnn_head_supported": (_I, [_I, _I, _I]),
    "mpnn_head_bn_linear_mse_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P]),
    "mpnn_head_bn_linear_mse_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "mpnn_bilinear_fwd": (_I, [_P, _P, _P, _L, _P, _L, _I, _P, _P]),
    "mpnn_bilinear_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _P, _I, _L, _I, _P, _P, _L, _P]),
    "mpnn_softmax_mul_fwd": (_I, [_P, _P, _L, _I, _P, _P, _P]),
    "mpnn_softmax_mul_bwd": (_I, [_P, _P, _P, _L, _I, _P, _P, _P]),
    "mpnn_dense_agg_fwd": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "mpnn_dense_agg_bwd": (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P]),
    "mpnn_prep_supported": (_I, [_I, _I, _I, _I]),
    "mpnn_prep_workspace_bytes": (_Z, [_I, _I]),
    "mpnn_prep_edges": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "mpnn_chain_supported": (_I, [_I, _I]),
    "mpnn_chain_debug": (_I, [_P]),
    "mpnn_chain_saved_floats": (_L, [_L, _I, _I]),
    "mpnn_chain_workspace_bytes": (_Z, [_L, _I, _I]),
    "mpnn_chain_fwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _PP, _I, _P, _P, _P, _P, _P, _P, _P, _P, _PP, _L, _I, _P,
                            _P, _P, _P, _Z, _P]),
    "mpnn_chain_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _PP, _I, _P, _P, _P, _P, _P, _P, _P, _P, _PP, _L, _I, _P,
                            _P, _P, _L, _P, _P, _P, _P, _P, _P, _PP, _P, _Z, _P]),
    "mpnn_real_rows_max": (_I, []),
    "mpnn_real_rows": (_I, [_P, _L, _P, _P, _Z, _P]),
    "mpnn_set2vec_set_persistent": (_I, [_I]),
    "mpnn_set2vec_debug": (None, [_P]),
    "mpnn_set2vec_saved_floats": (_L, [_I, _I, _I, _I]),
    "mpnn_set2vec_workspace_bytes": (_Z, [_I, _I, _I]),
    "mpnn_set2vec_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "mpnn_set2vec_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "mpnn_set2vec_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z,
                              _P]),
    "mpnn_lstm_hidden_fwd": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "mpnn_lstm_hidden_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
}

_lib = None


def load():
    """Loads libmpnn_b200.so (built in-tree by `python -m mpnn_b200.build`).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "mpnn_b200: %s not found -- run `python -m mpnn_b200.build` (there is no CPU/PyTorch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mpnn_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("mpnn_b200: internal error, non-contiguous tensor reached the C-ABI")
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (ctypes.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return ctypes.cast(arr, _PP)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = load().mpnn_last_error()
        raise RuntimeError("mpnn_b200.%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


_CLEAN_WS = {}


def clean_workspace(nbytes, device, family="bn"):
    """Persistent zero-initialised workspace per (op family, device, stream, size) for ops whose kernels leave their
    scratch counters / flags zero on exit (masked batch norms, the step kernels, the fused compaction): no memset per
    call.  Families never share a buffer: their flag words live at different offsets."""
    key = (family, device.index, torch.cuda.current_stream(device).cuda_stream, int(nbytes))
    ws = _CLEAN_WS.get(key)
    if ws is None:
        ws = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        _CLEAN_WS[key] = ws
    return ws


def f32c(t):
    """contiguous fp32 view/copy (the C-ABI contract)"""
    if t.dtype != torch.float32:
        raise RuntimeError("mpnn_b200: fp32 tensors only (got %s)" % t.dtype)
    return t.contiguous()
