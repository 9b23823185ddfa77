"""Warm per-kernel timeline of one graphed train step (CUPTI via torch.profiler): name, duration, gap to the
previous kernel.  Usage: python tools/timeline.py [config [batch [hidden]]] > gpurun_out/timeline.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "qm9"
    from mpnn_b200 import graph, synthetic
    from mpnn_b200.graphs import GraphedStep
    dev = torch.device("cuda:0")
    w = dict(bench.WORKLOADS[cfg])
    if len(sys.argv) > 2:
        w["B"] = int(sys.argv[2])
    if len(sys.argv) > 3:
        w["d"] = int(sys.argv[3])
        w["out"] = w["targets"] = 2 * w["d"]
    batch = bench.make_workload_batch(cfg, w, 0)
    devb = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask", "labels")}
    body, head = bench.build_model(w, dev)
    params = list(body.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, capturable=True, fused=True)

    def step(b):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        out = head(body(b["afm"], b["bfm"], b["adj"], b["mask"]))
        loss = torch.nn.functional.mse_loss(out, b["labels"])
        loss.backward()
        opt.step()
        return loss

    gs = GraphedStep(step, devb)
    for _ in range(5):
        gs.replay()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            gs.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    evs = evs[2 * n:]
    t0 = evs[0].time_range.start
    prev_end = t0
    tot_k = 0.0
    agg = {}
    for e in evs:
        dur = e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        prev_end = e.time_range.end
        tot_k += dur
        name = e.name.replace("(anonymous namespace)::", "")[:70]
        print("%8.1f  %6.1f  gap %5.1f  %s" % (e.time_range.start - t0, dur, gap, name))
        a = agg.setdefault(name[:50], [0, 0.0])
        a[0] += 1
        a[1] += dur
    span = evs[-1].time_range.end - t0
    print("kernels %d  span %.1f us  kernel time %.1f us  gaps %.1f us" % (len(evs), span, tot_k, span - tot_k))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("%-52s %3d %8.1f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / span))


if __name__ == "__main__":
    main()
