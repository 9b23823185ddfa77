"""mpnn_b200 -- B200-native (sm_100a) implementation of hochshi/mpnn's message-passing hot path.

Public surface: the reference's own plug-in API (`mpnn_b200.mpnn_functions`, `mpnn_b200.mask_batch_norm`,
see `mpnn_b200.dropin.install`) over a C-ABI CUDA library (`include/mpnn_b200.h`, `mpnn_b200/lib`).
"""
__version__ = "0.1.0"
