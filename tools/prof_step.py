"""One eager (un-graphed) train step of a bench workload, repeated a few times: the target of `ncu -k regex:...`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "qm9"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    from mpnn_b200 import graph
    dev = torch.device("cuda:0")
    w = bench.WORKLOADS[cfg]
    batch = bench.make_workload_batch(cfg, w, 0)
    devb = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask", "labels")}
    body, head = bench.build_model(w, dev)
    params = list(body.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    for _ in range(reps):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        out = head(body(devb["afm"], devb["bfm"], devb["adj"], devb["mask"]))
        loss = torch.nn.functional.mse_loss(out, devb["labels"])
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print("ok", float(loss))


if __name__ == "__main__":
    main()
