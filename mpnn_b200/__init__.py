"""mpnn_b200 -- B200-native (sm_100a) implementation of hochshi/mpnn's message-passing hot path.

Public surface: the reference's own plug-in API (`mpnn_b200.mpnn_functions`, `mpnn_b200.mask_batch_norm`,
see `mpnn_b200.dropin.install`) over a C-ABI CUDA library (`include/mpnn_b200.h`, `mpnn_b200/lib`).
"""
__version__ = "0.1.0"


def set_precision(mode):
    """"tf32" (default): feature widths 33..256 run on the tcgen05 kernels with TF32 operands (fp32 accumulate; forward
    within SURVEY 8c's 2e-2 after the GRU / readout, gradients of saturated GRU units can be further off when the
    messages are large).  "fp32": every width on the fp32 kernels -- the reference's accuracy (1e-4 / 1e-3), slower.
    Process-wide; returns the previous mode."""
    from . import _lib
    if mode not in ("tf32", "fp32"):
        raise ValueError("mpnn_b200.set_precision: 'tf32' or 'fp32'")
    prev = _lib.load().mpnn_set_tensor_cores(1 if mode == "tf32" else 0)
    return "tf32" if prev else "fp32"
