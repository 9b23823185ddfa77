"""Pinned host -> device and device -> device copy times at the sizes of BASELINE config 2's batch (0.64 MB ragged, 7.4 MB
padded) and a large one: is the e2e step bound by PCIe?  (Measured on the pool's B200 boxes: 7.4 MB H2D in 140 us = 53 GB/s,
shorter than the 0.22 ms step it hides behind.)
    python tools/h2d_probe.py"""
import torch
dev = torch.device("cuda:0")
for mb in (0.64, 7.4, 64.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2D", lambda: d2.copy_(d, non_blocking=True))):
        ts = []
        for i in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print("%s %.2f MB: median %.1f us = %.1f GB/s" % (name, mb, ts[10], n / ts[10] / 1e3))
