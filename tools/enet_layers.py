"""Per-layer cost of the fused edge-network kernels (csrc/typed.cu k_enet_fwd / k_enet_bwd): times the forward and the
backward of one network on R distinct rows for several numbers of tied layers; the slope is the cost of one layer of
the serial chain, the intercept everything else (launch, growth layers, table / last-layer gradient).
    python tools/enet_layers.py [R]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpnn_b200 import _lib
from mpnn_b200.functional import EdgeNetTableFn, MultiEdgeNetTableFn


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ef, P, nf = 7, 49, 16
    urows = torch.rand(R, ef, device=dev)
    w_tied = (torch.randn(P, P, device=dev) * (2.0 / P) ** 0.5).requires_grad_(True)
    W_last = (torch.randn(nf * nf, P, device=dev) * 0.1).requires_grad_(True)
    B_last = torch.zeros(nf * nf, device=dev, requires_grad=True)
    gw = (torch.randn(P, ef, device=dev) * 0.3).requires_grad_(True)
    gb = torch.zeros(P, device=dev, requires_grad=True)
    params = [w_tied, W_last, B_last, gw, gb]
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    multi = []
    for k in range(K):
        multi += [(p_.detach().clone() + 0.01 * k).requires_grad_(True) for p_ in params]

    def fwd(L):
        if K > 1:    # K sibling networks in one launch each way (one EdgeNetwork per message-passing step)
            outs = MultiEdgeNetTableFn.apply(urows, L, nf, nf, K, 1, *multi)
            return torch.stack([outs[2 * k] for k in range(K)]).sum(0)
        return EdgeNetTableFn.apply(urows, w_tied, L, W_last, B_last, nf, nf, gw, gb)[0]

    def both(L):
        for p_ in params + multi:
            p_.grad = None
        t = fwd(L)
        t.backward(gones)

    gones = torch.ones(R, 16, 16, device=dev)
    for L in (2, 10, 26, 50):
        res = []
        for fn in (fwd, both):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn(L)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn(L)
            ts = []
            for it in range(30):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            res.append(ts[len(ts) // 2])
        print("L=%2d  fwd %.1f us   fwd + bwd + finish %.1f us   (difference %.1f us)" % (L, res[0], res[1], res[1] - res[0]))


if __name__ == "__main__":
    main()
