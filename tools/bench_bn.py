"""Isolated timing of MaskBatchNorm forward/backward (csrc/bn.cu), 50 calls replayed from a CUDA graph.
Usage: python tools/bench_bn.py [rows C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpnn_b200 import _lib
from mpnn_b200._lib import ptr, check

def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 7424
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    lib = _lib.load()
    dev = torch.device("cuda:0")
    x = torch.randn(rows, C, device=dev)
    mask = (torch.rand(rows, device=dev) > 0.4).float()
    x = x * mask[:, None]
    y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    stats = torch.empty(2 * C + 1, device=dev)
    ws = torch.zeros(lib.mpnn_bn_workspace_bytes(rows, C), dtype=torch.uint8, device=dev)
    def fwd():
        check(lib.mpnn_mask_bn_fwd(ptr(x), ptr(mask), rows, C, 1e-6, ptr(y), ptr(stats), ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream), "f")
    def bwd():
        check(lib.mpnn_mask_bn_bwd(ptr(x), ptr(mask), ptr(dy), ptr(stats), rows, C, ptr(dx), ptr(ws), ws.numel(), torch.cuda.current_stream().cuda_stream), "b")
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(5):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        print("%s rows=%d C=%d: %.2f us per call (graph of 50)" % (name, rows, C, a.elapsed_time(b) * 1000 / 50))

main()
