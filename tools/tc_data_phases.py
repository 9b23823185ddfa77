"""%globaltimer stamps of k_tc_gru_data_grad's producer (CTA 0) at B = 16 384, d = 64: where a group's time goes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpnn_b200 import _lib, functional as Fn

dev = torch.device("cuda:0")
lib = _lib.load()
rows, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (377517, 64)
g = torch.Generator().manual_seed(0)
m, h = torch.randn(rows, d, generator=g).to(dev), torch.randn(rows, d, generator=g).to(dev)
mask = torch.ones(rows, device=dev)
W = [(torch.randn(d, 3 * d, generator=g) * 0.2).to(dev).requires_grad_(True) for _ in range(2)]
bb = [torch.zeros(3 * d, device=dev, requires_grad=True) for _ in range(2)]
mm, hh = m.requires_grad_(True), h.requires_grad_(True)
out = Fn.GRUFn.apply(mm, hh, mask, W[0], W[1], bb[0], bb[1], None)
cot = torch.randn_like(out)
for it in range(3):
    dbg = torch.zeros(64, dtype=torch.int64, device=dev)
    lib.mpnn_tc_debug(ctypes.c_void_p(dbg.data_ptr()))
    torch.autograd.grad(out, [mm, hh] + W + bb, cot, retain_graph=True)
    torch.cuda.synchronize()
    lib.mpnn_tc_debug(None)
    v = dbg.cpu().tolist()
    names = ["stages free", "half 0 landed", "half 1 landed (h0 converted)", "converted", "pass-1 stages free", "copied"]
    for gi in range(2, 6):
        t = v[gi * 6:(gi + 1) * 6 + 1]
        if all(t):
            print("group %d (us): " % gi + ", ".join("%s %.2f" % (names[k], (t[k + 1] - t[k]) / 1e3) for k in range(6)),
                  "| total %.2f" % ((t[6] - t[0]) / 1e3))
