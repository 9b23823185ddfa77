"""Phase timestamps of CTA 0 of the fused forward step kernel (MPNN_B200_CHAIN_DEBUG=1): where its time goes."""
import ctypes, os, sys
os.environ["MPNN_B200_CHAIN_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mpnn_b200 import _lib, graph

dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "qm9"
w = dict(bench.WORKLOADS[cfg])
batch = bench.make_workload_batch(cfg, w, 0)
t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
body, head = bench.build_model(w, dev)
lib = _lib.load()
for it in range(4):
    graph.clear_cache()
    out = body(t["afm"], t["bfm"], t["adj"], t["mask"])
    out.sum().backward()
    buf = (ctypes.c_longlong * (64 + 3 * 640))()
    lib.mpnn_chain_debug(buf)
    v = list(buf)
    n = max(i for i in range(40) if v[i]) + 1
    print("iter", it, " ".join("%.1f" % ((v[i] - v[0]) / 1e3) for i in range(n)))
    print("   last arriver per barrier (arrive, combined, cta):", [(round((v[40 + 3 * k] - v[0]) / 1e3, 1), round((v[41 + 3 * k] - v[0]) / 1e3, 1), v[42 + 3 * k]) for k in range(3)])
