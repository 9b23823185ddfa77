"""Per-kernel totals of one EAGER train step of a bench workload (CUPTI via torch.profiler).
Usage: python tools/prof_eager.py [config]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "zinc"
    from mpnn_b200 import graph
    dev = torch.device("cuda:0")
    w = dict(bench.WORKLOADS[cfg])
    batch = bench.make_workload_batch(cfg, w, 0)
    devb = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask", "labels")}
    body, head = bench.build_model(w, dev)
    params = list(body.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)

    def step():
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        out = head(body(devb["afm"], devb["bfm"], devb["adj"], devb["mask"]))
        loss = torch.nn.functional.mse_loss(out, devb["labels"])
        loss.backward()
        opt.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    span = evs[-1].time_range.end - evs[0].time_range.start
    agg = {}
    for e in evs:
        a = agg.setdefault(e.name.replace("(anonymous namespace)::", "")[:60], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    tot = sum(v[1] for v in agg.values())
    print("kernels %d  span %.1f us  kernel time %.1f us" % (len(evs), span, tot))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%-60s %5d %9.1f %5.1f%%" % (k, c, us, 100 * us / tot))


if __name__ == "__main__":
    main()
