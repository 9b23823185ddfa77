"""Set2Vec alone (BASELINE config 3's readout: B=128, N=38, F=64, 100 iterations): persistent kernels vs per-iteration
launches, CUDA-event timed, plus the %globaltimer phase stamps of CTA 0 (mpnn_set2vec_debug).

    python tools/s2v_bench.py [B N nf steps]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpnn_b200 import _lib, functional, modules as M   # noqa: E402

FWD = ["lstm", "query", "energy", "stats+post", "read+gemvH", "gather", "finalize", "gemvR"]
BWD = ["datt", "gather", "de+energy", "reduce+dh", "lstm", "dm_prev"]


def main():
    B, N, nf, steps = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (128, 38, 32, 100)
    dev = torch.device("cuda:0")
    lib = _lib.load()
    torch.manual_seed(0)
    mod = M.Set2Vec(nf, 99, time_steps=steps).to(dev)
    X = torch.randn(B, N, 2 * nf, device=dev)
    n = torch.randint(N // 2, N + 1, (B,), device=dev)
    mask = (torch.arange(N, device=dev)[None, :] < n[:, None]).float()[..., None]
    X = X * mask

    def step():
        x = X.clone().requires_grad_(True)
        out = mod(x, mask)
        out.sum().backward()

    for persistent in (1, 0):
        functional.set2vec_persistent(persistent)
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(10):
            x = X.clone().requires_grad_(True)
            ev[0].record()
            out = mod(x, mask)
            ev[1].record()
            out.sum().backward()
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        print("persistent=%d  B=%d N=%d F=%d steps=%d: forward %.3f ms, backward %.3f ms" %
              (persistent, B, N, 2 * nf, steps, tf / 10, tb / 10))
    functional.set2vec_persistent(1)
    functional.OP_GRAPHS_ENABLED = False
    dbg = torch.zeros(448, dtype=torch.int64, device=dev)
    lib.mpnn_set2vec_debug(ctypes.c_void_p(dbg.data_ptr()))
    step()
    torch.cuda.synchronize()
    lib.mpnn_set2vec_debug(None)
    d = dbg.cpu().tolist()
    for base, names, label in ((0, FWD, "forward"), (64, BWD, "backward")):
        for it in range(1, 4):
            t = d[base + it * 16: base + it * 16 + len(names) + 1]
            nxt = d[base + (it + 1) * 16] if it < 3 else None
            parts = ["%s %.2f" % (names[k], (t[k + 1] - t[k]) / 1e3) for k in range(len(names)) if t[k + 1] and t[k]]
            tot = (nxt - t[0]) / 1e3 if nxt else float("nan")
            print("%s iteration %d (us): %s | iteration %.2f" % (label, it, ", ".join(parts), tot))
    skew(d)



def skew(d):
    post = [v for v in d[128::2] if v]
    done = [v for v in d[129::2] if v]
    if post and done:
        t0 = min(post)
        print("forward iteration 2, all CTAs (us after the first post): posts %.2f .. %.2f, gather done %.2f .. %.2f" %
              (0.0, (max(post) - t0) / 1e3, (min(done) - t0) / 1e3, (max(done) - t0) / 1e3))


if __name__ == "__main__":
    main()
