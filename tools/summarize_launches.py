"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: splits it into train steps (each step
starts with the compaction kernel `k_count`) and prints the per-kernel table of one step.
    python tools/summarize_launches.py gpurun_out/launches.csv [step_index]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    pick = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        rows.append((r[ki], us))
    starts = [i for i, (k, _) in enumerate(rows) if "k_count(" in k]
    steps = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]
    print("launches %d, steps found %d (lengths %s)" % (len(rows), len(steps), sorted(set(len(s) for s in steps))))
    step = steps[pick]
    tot = sum(us for _, us in step)
    agg = collections.OrderedDict()
    for k, us in step:
        k = k.replace("<unnamed>::", "").replace("void ", "").split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    print("step %d: %d launches, sum of launch durations %.1f us" % (pick, len(step), tot))
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, c, us, 100 * us / tot))


if __name__ == "__main__":
    main()
