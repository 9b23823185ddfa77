"""Kernel-level sweep of the tcgen05 typed message path (csrc/tc_message.cu) at BASELINE config-5 sizes:
ZINC-shaped graphs, B in {512..16384}, hidden 64/128/256.  Times, with CUDA events and an L2 flush between
repetitions, the grouped edge GEMM (forward / d-sender form), the CSR segmented sum and the table gradient, and
prints one JSON line per (B, d) with achieved TFLOP/s and GB/s against MEASURED_PEAKS.json.

    python tools/bench_tc.py [--B 512,4096,16384] [--d 64,128,256] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, flush, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", default="512,4096,16384")
    ap.add_argument("--d", default="64,128,256")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    from mpnn_b200 import _lib, graph, synthetic
    from mpnn_b200._lib import check, ptr, stream
    lib = _lib.load()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for B in [int(x) for x in args.B.split(",")]:
        b = synthetic.make_batch("autoenc", B=B)
        bfm, adj = torch.from_numpy(b["bfm"]).to(dev), torch.from_numpy(b["adj"]).to(dev)
        graph.clear_cache()
        el = graph.compact_edges(bfm, adj)
        del bfm, adj
        ti = el.typed()
        plan = ti.tc_plan(el)
        E, n_rows, n = el.E, el.n_rows, int(b["n_atoms"])
        for d in [int(x) for x in args.d.split(",")]:
            DP = lib.mpnn_tc_dp(d, d)
            g = torch.Generator().manual_seed(d)
            H = torch.randn(n_rows, d, generator=g).to(dev)
            dM = torch.randn(n_rows, d, generator=g).to(dev)
            table = torch.randn(ti.Ucap + 1, DP, DP, generator=g).to(dev)
            Y = torch.empty(E, d, device=dev)
            M = torch.empty(n_rows, d, device=dev)
            dT = torch.empty_like(table)
            ws = _lib.workspace(lib.mpnn_tc_table_grad_workspace_bytes(ti.Ucap, DP), dev)
            ws2 = _lib.workspace(lib.mpnn_tc_edge_gemm_workspace_bytes(ti.Ucap, DP), dev)

            def gemm():
                check(lib.mpnn_tc_edge_gemm(ptr(plan), el.Ecap, ti.Ucap, ptr(ti.type_eid), 0, ptr(H), d, d, ptr(table), DP,
                                            1, ptr(Y), d, d, ptr(ws2), ws2.numel(), stream()), "gemm")

            def seg():
                check(lib.mpnn_segment_sum(ptr(Y), ptr(el.row_ptr), None, n_rows, d, d, ptr(M), d, 0, 1.0, stream()), "seg")

            def grad():
                check(lib.mpnn_tc_table_grad(ptr(plan), el.Ecap, ti.Ucap, ptr(H), d, ptr(dM), d, DP, 1, ptr(dT), ptr(ws),
                                             ws.numel(), stream()), "grad")

            t_gemm, t_seg, t_grad = timed(gemm, flush, args.reps), timed(seg, flush, args.reps), timed(grad, flush, args.reps)
            # masked GRU forward (fused tcgen05 kernel for d <= 128, dense GEMMs + pointwise above)
            mask = (torch.rand(n_rows, generator=g) > 0.4).float().to(dev)
            W_ih = (torch.randn(d, 3 * d, generator=g) / d ** 0.5).to(dev)
            W_hh = (torch.randn(d, 3 * d, generator=g) / d ** 0.5).to(dev)
            b3 = torch.zeros(3 * d, device=dev)
            h_out = torch.empty(n_rows, d, device=dev)
            gates = torch.empty(n_rows, 4 * d, device=dev)
            wsg = _lib.workspace(lib.mpnn_gru_workspace_bytes(n_rows, d), dev)

            def gru():
                check(lib.mpnn_gru_fwd(ptr(M), ptr(H), ptr(mask), ptr(W_ih), ptr(W_hh), ptr(b3), ptr(b3), n_rows, d,
                                       ptr(h_out), ptr(gates), ptr(wsg), wsg.numel(), stream()), "gru")

            t_gru = timed(gru, flush, args.reps)
            gru_bytes = 4.0 * n_rows * d * 8 + 4.0 * n_rows     # m, h in; h' and the 4 saved gate planes out
            gru_flops = 12.0 * n_rows * d * d
            flops = 2.0 * E * d * d
            # algorithmic bytes: every edge reads one sender row and writes one message row (+ 3 index/weight words)
            gemm_bytes = 4.0 * E * d * 2 + 12.0 * E + 4.0 * (ti.U + 1) * d * d
            seg_bytes = 4.0 * E * d + 4.0 * n_rows * d + 4.0 * n_rows
            grad_bytes = 4.0 * E * d * 2 + 16.0 * E + 4.0 * (ti.U + 1) * d * d
            line = {
                "B": B, "d": d, "DP": DP, "atoms": n, "edges": E, "types": ti.U,
                "edge_gemm": {"ms": t_gemm, "tflops": flops / t_gemm / 1e9, "gbs": gemm_bytes / t_gemm / 1e6,
                              "frac_hbm": gemm_bytes / t_gemm / 1e6 / peaks["hbm_gbs"],
                              "frac_tensor_bf16": flops / t_gemm / 1e9 / peaks["bf16_tflops"]},
                "segment_sum": {"ms": t_seg, "gbs": seg_bytes / t_seg / 1e6, "frac_hbm": seg_bytes / t_seg / 1e6 / peaks["hbm_gbs"]},
                "table_grad": {"ms": t_grad, "tflops": flops / t_grad / 1e9, "gbs": grad_bytes / t_grad / 1e6,
                               "frac_hbm": grad_bytes / t_grad / 1e6 / peaks["hbm_gbs"]},
                "gru_fwd": {"ms": t_gru, "tflops": gru_flops / t_gru / 1e9, "gbs": gru_bytes / t_gru / 1e6,
                            "frac_hbm": gru_bytes / t_gru / 1e6 / peaks["hbm_gbs"]},
                # one forward message-passing step = message GEMM + aggregation + GRU update
                "mp_step_fwd": {"ms": t_gemm + t_seg + t_gru,
                                "gbs": (gemm_bytes + seg_bytes + gru_bytes) / (t_gemm + t_seg + t_gru) / 1e6,
                                "frac_hbm": (gemm_bytes + seg_bytes + gru_bytes) / (t_gemm + t_seg + t_gru) / 1e6 / peaks["hbm_gbs"]},
            }
            print(json.dumps(line), flush=True)
            del H, dM, table, Y, M, dT, ws, gates, h_out, wsg


if __name__ == "__main__":
    main()
