"""bench.py -- graphs/sec of one fwd+bwd training step of the message-passing path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config qm9|lipo|autoenc] [--impl ours|reference]

Workload (N=1 default): BASELINE.json configs[1] -- normed_basic_model (one EdgeNetwork per step, AdjMsgAgg,
masked GRU, MaskBatchNorm, GraphLevelOutput) on QM9-shaped synthetic graphs, batch 256 per GPU, d=16, ef=7,
P=49, T=3, 12 regression targets, MSE + Adam.  A "step" = forward + loss + backward + optimizer step
(+ gradient all-reduce when N>1; weak scaling: 256 graphs per GPU).

JSON keys beyond the base contract:
  roofline     SURVEY 8d's per-message-passing-step figure for the headline workload: the step kernels alone (message
               function + aggregation + GRU + masked BN, forward and backward), replayed from CUDA graphs with the L2
               flushed, algorithmic fwd+bwd work / time against the MEASURED peaks (MEASURED_PEAKS.json)
  roofline_large  the same at sizes where HBM is the limit: basic_graph_autoencoder, B=16384, d=64 and d=256
  median_ms_per_step / median_value  median over the timed steps (value is steps / total time, as the contract asks)
  cpu_baseline the oracle port of the reference's CPU path (oracle/mpnn_oracle.py), timed on this box's host
               cores on a bounded sample of the same workload
  e2e          the same metric through the public module API with HOST (pinned) inputs: H2D of the step's
               padded batch and a D2H read of the loss inside the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "graphs/sec (fwd+bwd train step)"

WORKLOADS = {
    # name: (synthetic config, caller variant, d, ef, T, readout width, targets)
    "qm9": dict(variant="normed", d=16, ef=7, T=3, out=64, targets=12, B=256, head="bn_linear", pipeline_prep=True,
                desc="normed_basic_model + MaskBatchNorm + BatchNorm1d(64) + Linear(64,12), QM9-shaped (n<=29), "
                     "B=256/GPU, d=16, ef=7, P=49, T=3"),
    # the same model on NON-categorical bond features (the reference's docstrings allow "topological distance and 3D
    # distance", models/basic_model.py:41-42): every bond row is distinct, the edge network runs per edge (trunk on E rows,
    # per-edge contraction kernels, csrc/trunk.cu + csrc/message.cu); eager launches (the per-edge path sizes its arrays
    # from the host-side edge count)
    "qm9_continuous": dict(variant="normed", d=16, ef=7, T=3, out=64, targets=12, B=256, head="bn_linear", synth="qm9",
                           continuous=True, graph=False,
                           desc="normed_basic_model as config 2 but with a continuous bond feature (one distinct row per "
                                "bond): per-edge contraction path, eager launches, B=256/GPU, d=16, ef=7, P=49, T=3"),
    "lipo": dict(variant="lipo", d=19, ef=7, T=6, out=38, targets=1, B=32, head="bn_halving",
                 desc="lipo_basic_model (HEAD form) + BatchNorm1d(38) + halving dense head, Lipophilicity-shaped, B=32, "
                      "d=19, ef=7, P=49, T=6"),
    "autoenc": dict(variant="autoencoder", d=64, ef=8, T=3, out=128, targets=128, B=512,
                    desc="basic_graph_autoencoder.encode, ZINC-shaped, B=512/GPU, d=64, ef=8, P=64, T=3"),
    # configs[2]: att_model (AttEdgeNetwork + AdjMsgAgg + MaskBatchNorm + Set2Vec, 100 steps), 128 graphs per GPU
    "zinc": dict(variant="att", d=32, ef=8, T=3, out=128, targets=1, B=128, message_func="AttEdgeNetwork",
                 readout_func="Set2Vec",
                 desc="att_model (AttEdgeNetwork, AdjMsgAgg, Set2Vec x100), ZINC-shaped, B=128/GPU, d=32, ef=8, P=64, T=3"),
    # configs[3]: normed_encoded_basic_model (atom/bond encoders + masked BN1d everywhere), B=2048 global
    "affinity": dict(variant="normed_encoded", d=8, ef=2, T=3, out=16, targets=1, B=2048, encoders=True,
                     desc="normed_encoded_basic_model (encoders 30->8 / 8->2, MaskBatchNorm1d), B=2048, d=8, ef=2, P=16, T=3"),
    # configs[3] as BASELINE.json names it: normed_encoded_basic_model_ecfp (no ma_bn, obn on the per-atom readout the model
    # was written against, graph_level_output.py:46), per-atom targets like the driver's masked ECFP regression
    # (test_graph_encode_norm_ecfp.py:137), B=2048
    "affinity_ecfp": dict(variant="normed_encoded_ecfp", d=8, ef=2, T=3, out=16, targets=1, B=2048, encoders=True,
                          readout_func="GraphLevelOutputAtoms", per_atom=True, synth="affinity",
                          desc="normed_encoded_basic_model_ecfp (encoders 30->8 / 8->2, MaskBatchNorm1d per step + obn, "
                               "per-atom readout), B=2048, d=8, ef=2, P=16, T=3"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, gpu_index):
        super(ClockSampler, self).__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_workload_batch(config, w, rank):
    from mpnn_b200 import synthetic
    batch = synthetic.make_batch(w.get("synth", config), B=w["B"], seed_offset=rank,
                                 d=w["d"] if config == "autoenc" else None, return_graphs=True)
    if w.get("continuous"):     # a "3D distance"-like column: symmetric, positive on bonds, zero elsewhere
        rs = np.random.RandomState(1000 + rank)
        dist = rs.uniform(0.9, 1.6, size=batch["adj"].shape).astype(np.float32)
        dist = np.maximum(dist, dist.transpose(0, 2, 1)) * (batch["adj"] != 0)
        batch["bfm"] = batch["bfm"].copy()
        batch["bfm"][..., -1] = dist
    if config not in ("qm9", "qm9_continuous"):
        shape = (w["B"], batch["afm"].shape[1], w["targets"]) if w.get("per_atom") else (w["B"], w["targets"])
        batch["labels"] = np.random.RandomState(rank).normal(size=shape).astype(np.float32)
    return batch


def build_model(w, dev):
    # the UNMODIFIED reference model files (tests/ref_models/models/*.py, byte-identical) over this package's modules
    from mpnn_b200.dropin import reference_model as MessagePassingModel, kaiming_init
    from mpnn_b200 import modules as M
    torch.manual_seed(317)
    kw = {}
    if w.get("message_func"):
        kw["message_func"] = getattr(M, w["message_func"])
    if w.get("readout_func"):
        kw["readout_func"] = getattr(M, w["readout_func"])
    if w.get("encoders"):   # AtomAutoEncoder / BondAutoEncoder .encoder halves (encoders/*_autoencoder.py:7-11)
        kw["atom_encoder"] = M.AtomAutoEncoder().encoder
        kw["bond_encoder"] = M.BondAutoEncoder().encoder
    body = MessagePassingModel(w["variant"], w["d"], w["ef"], w["d"], 1, w["out"], message_steps=w["T"], **kw)
    body.apply(kaiming_init)
    # prediction heads of the reference drivers (stock torch modules): test_graph_norm.py:86-90
    # BatchNorm1d(out) + Linear(out, targets); test_lipo.py:103-129 BatchNorm1d(out) + halving dense stack
    if w.get("head") == "bn_linear":
        head = torch.nn.Sequential(torch.nn.BatchNorm1d(w["out"]), torch.nn.Linear(w["out"], w["targets"]))
    elif w.get("head") == "bn_halving":
        layers, den = [torch.nn.BatchNorm1d(w["out"])], w["out"]
        while den > 10:
            nd = int(np.ceil(den / 2))
            layers += [torch.nn.Linear(den, nd), torch.nn.ReLU()]
            den = nd
        layers.append(torch.nn.Linear(den, w["targets"]))
        head = torch.nn.Sequential(*layers)
    else:
        head = torch.nn.Linear(w["out"], w["targets"])
    return body.to(dev), head.to(dev)


def run_ours(args):
    from mpnn_b200 import _lib, dist as D, graph, synthetic
    _lib.load()  # fail loudly if the CUDA library is missing
    rank, world = D.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = dict(WORKLOADS[args.config])
    if args.batch:
        w["B"] = args.batch
        w["desc"] += " [batch overridden: %d]" % args.batch
    if args.hidden:
        w["d"] = args.hidden
        w["out"] = w["targets"] = 2 * args.hidden
        w["desc"] += " [hidden overridden: %d]" % args.hidden
    B = w["B"]
    batch = make_workload_batch(args.config, w, rank)
    n, e = batch["n_atoms"], batch["n_edges"]
    keys = ("afm", "bfm", "adj", "mask", "labels")
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in keys}
    devb = {k: v.to(dev) for k, v in host.items()}
    body, head = build_model(w, dev)
    params = list(body.parameters()) + list(head.parameters())
    # the attention gate (per-pair sender vectors) and differentiable bond features run on the per-edge contraction
    # kernels, whose array sizes come from a host read of the edge count: eager launches for those workloads
    use_graph = not args.no_graph and w.get("graph", True)
    if args.stock_adam:
        opt = torch.optim.Adam(params, lr=1e-3, capturable=use_graph, fused=use_graph or None)
    else:   # same update rule as torch.optim.Adam, one launch over all parameter tensors (mpnn_b200/optim.py)
        from mpnn_b200.optim import FusedAdam
        opt = FusedAdam(params, lr=1e-3)
    allreduce = D.FlatGradAllReduce(params)
    fused_ddp = False
    if world > 1 and not args.stock_adam and not args.nccl_allreduce:
        # the gradient all-reduce runs INSIDE the Adam launch over NVLink peer memory (mpnn_b200/optim.py enable_ddp)
        fused_ddp = opt.enable_ddp()
    # L2 (126 MB) is flushed between timed iterations by writing a 256 MB buffer
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # the driver's head + criterion (BatchNorm1d -> Linear -> MSELoss) as one launch each way (mpnn_b200/heads.py wraps
    # the same stock modules; --stock-head keeps them as 14 torch kernels)
    fused_head = None
    if w.get("head") == "bn_linear" and not args.stock_head:
        from mpnn_b200.heads import BNLinearMSE
        fused_head = BNLinearMSE(head[0], head[1])

    def step(b):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        feats = body(b["afm"], b["bfm"], b["adj"], b["mask"])
        if fused_head is not None:
            loss = fused_head(feats, b["labels"])
        else:
            loss = torch.nn.functional.mse_loss(head(feats), b["labels"])
        loss.backward(gradient=one)     # (a persistent 1.0: no fill launch for the seed of the backward pass)
        if not fused_ddp:
            allreduce()
        opt.step()
        return loss

    one = torch.ones((), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    gs = gs_io = None
    if use_graph:
        # the whole step (compaction .. optimizer) as ONE CUDA graph over static input buffers
        from mpnn_b200.graphs import GraphedStep
        # input preprocessing (compaction, de-duplication, type sort: functions of bfm / adj only) software-pipelined one
        # step ahead, as a branch of the previous step's graph, where the one-launch compaction serves the batch shape
        pipe = (not args.no_pipeline_prep and w.get("pipeline_prep", False)
                and bool(_lib.load().mpnn_prep_supported(devb["bfm"].shape[0], devb["bfm"].shape[1],
                                                         devb["bfm"].shape[3], 64)))
        gs = GraphedStep(step, devb, warmup=3, pipeline_prep=pipe)
        # e2e staging: every group of inputs ("bonds" = bfm / adj, "rest" = afm / mask / labels) is one flat allocation on
        # the host (pinned), in the device staging area and in the graph's static inputs: two copies move a batch.
        # Pipelined prep: a prefetched batch's bfm / adj are consumed one replay before its afm / mask / labels, which wait
        # in one of two small staging sets (alternating by step).
        rest_keys = [k for k in keys if k not in ("bfm", "adj")]
        host_flat = {}
        for g in ("bonds", "rest"):
            host_flat[g], hv = gs.mirror(g, device="cpu", pin_memory=True)
            for k, v in hv.items():
                v.copy_(host[k])
        small = []
        for _ in range(2):
            f, v = gs.mirror("rest", device=dev)
            f.copy_(host_flat["rest"])
            small.append((f, v))
        staging_bonds, _ = gs.mirror("bonds", device=dev)
        staging_rest, _ = gs.mirror("rest", device=dev)
        cnt = {"pre": 0, "run": 0, "run_r": 0}

        def run_resident():
            return gs.replay()

        # e2e input pipeline: like a prefetching DataLoader (pin_memory + non_blocking), the NEXT step's host->device
        # copy runs on a copy stream into a staging buffer while the current step computes; every timed step still
        # contains one full H2D of a padded batch, a D2D into the graph's static inputs and a synchronous loss read.
        copy_stream = torch.cuda.Stream()
        ev_copy, ev_loaded = torch.cuda.Event(), torch.cuda.Event()

        def prefetch():
            with torch.cuda.stream(copy_stream):
                j = cnt["pre"]
                cnt["pre"] = j + 1
                staging_bonds.copy_(host_flat["bonds"], non_blocking=True)
                (small[j & 1][0] if pipe else staging_rest).copy_(host_flat["rest"], non_blocking=True)
                ev_copy.record(copy_stream)

        def run_e2e():
            main = torch.cuda.current_stream()
            main.wait_event(ev_copy)           # this step's inputs have landed in the staging buffer
            if pipe:
                # the batch that has just landed is the one AFTER the batch this replay trains on: its bfm / adj are
                # prepared during this replay; the current batch's afm / mask / labels arrived one step earlier
                k_ = cnt["run"]
                cnt["run"] = k_ + 1
                gs.flat["rest"].copy_(small[(k_ - 1) & 1][0], non_blocking=True)
            else:
                gs.flat["rest"].copy_(staging_rest, non_blocking=True)
            gs.flat["bonds"].copy_(staging_bonds, non_blocking=True)
            ev_loaded.record(main)
            copy_stream.wait_event(ev_loaded)  # the staging buffer may be overwritten from here on
            prefetch()                         # next step's H2D overlaps this step's kernels
            return gs.replay()

        gs_io = None
        if not args.no_host_io:
            # the same pipeline as ONE graph per step: device copy staging -> static inputs at the head, the step, and the
            # H2D of the host's pinned buffers into the staging area as a parallel branch (GraphedStep(host_io=True));
            # the host launches one graph and reads the previous step's loss
            gs_io = GraphedStep(step, devb, warmup=1, pipeline_prep=pipe, host_io=True)
            run_e2e_staged, prefetch_staged = run_e2e, prefetch

            def run_e2e():
                return gs_io.replay()

            def prefetch():
                pass
    else:
        def run_resident():
            return step(devb)

        def prefetch():
            pass

        def run_e2e():
            return step({k: v.to(dev, non_blocking=True) for k, v in host.items()})
    for _ in range(args.warmup):
        run_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- value: inputs resident in HBM -------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 1)
        ev[i][0].record()
        run_resident()
        ev[i][1].record()
    barrier()
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms = sum(per_step)
    med_local = float(np.median(per_step))
    # ---- e2e: host buffers, H2D + loss D2H inside the timed region ----------------------------
    def time_e2e(run_fn, prefetch_fn):
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        prefetch_fn()   # the first step's inputs; every timed step issues the copy for its successor
        barrier()
        # every step's loss is copied to pinned host memory inside the step's timed region and READ by the host one
        # step later (asynchronous logging): the host never stalls the GPU between steps
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        losses, prev = [], None
        for i in range(args.steps):
            flush.fill_(i & 1)
            ev2[i][0].record()
            loss = run_fn()
            loss_host[i & 1].copy_(loss.detach().reshape(()), non_blocking=True)
            loss_ev[i & 1].record()
            if prev is not None:
                loss_ev[prev].synchronize()
                losses.append(float(loss_host[prev]))
            ev2[i][1].record()
            prev = i & 1
        loss_ev[prev].synchronize()
        losses.append(float(loss_host[prev]))
        barrier()
        assert len(losses) == args.steps and all(np.isfinite(losses)), "e2e: a step's loss was not read back"
        return sum(a.elapsed_time(b) for a, b in ev2)

    ms2 = time_e2e(run_e2e, prefetch)
    # ---- e2e through the device-side collate (mpnn_b200.loader): the host ships the batch ragged
    # (real atoms' rows + edge list) and the padded tensors the modules consume are written on the GPU
    ms3, rb_bytes = None, None
    if use_graph:
        from mpnn_b200.loader import RaggedBatch
        rb_host = RaggedBatch.from_graphs(batch["graphs"], batch["labels"])
        rb_bytes = rb_host.nbytes()
        rb_dev = rb_host.to(dev, non_blocking=False)       # device staging buffers of the ragged arrays
        pad_out = {k: gs.static[k] for k in ("afm", "bfm", "adj", "mask")}

        def prefetch_r():
            with torch.cuda.stream(copy_stream):
                for k, v in rb_host.tensors().items():
                    rb_dev.tensors()[k].copy_(v, non_blocking=True)
                ev_copy.record(copy_stream)

        def run_e2e_r():
            main = torch.cuda.current_stream()
            main.wait_event(ev_copy)
            if pipe:
                # the ragged batch that has landed is the NEXT one: bfm / adj straight into the graph's inputs, afm / mask
                # / labels into the small set the next step loads from
                k_ = cnt["run_r"]
                cnt["run_r"] = k_ + 1
                gs.flat["rest"].copy_(small[(k_ - 1) & 1][0], non_blocking=True)
                nxt = small[k_ & 1][1]
                rb_dev.scatter_padded({"afm": nxt["afm"], "bfm": gs.static["bfm"], "adj": gs.static["adj"],
                                       "mask": nxt["mask"]})
                nxt["labels"].copy_(rb_dev.labels, non_blocking=True)
            else:
                rb_dev.scatter_padded(pad_out)              # memsets + two scatter kernels into the graph's inputs
                gs.static["labels"].copy_(rb_dev.labels, non_blocking=True)
            ev_loaded.record(main)
            copy_stream.wait_event(ev_loaded)
            prefetch_r()
            return gs.replay()

        for _ in range(2):
            prefetch_r()
            run_e2e_r()
        ms3 = time_e2e(run_e2e_r, prefetch_r)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms2, ms3 if ms3 is not None else 0.0, med_local], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms2 = float(t[0]), float(t[1])
    med_ms = float(t[3])
    ms3 = float(t[2]) if ms3 is not None else None

    # ---- roofline of the dominant kernel, timed alone with CUDA events (L2 flushed before every launch) -------
    roof = None
    roof_large = None
    launches = None
    if rank == 0 and w["variant"] in ("normed", "basic", "autoencoder") and not args.no_roofline:
        net0 = body.mfs[0] if hasattr(body, "mfs") else body.mf
        roof = mp_step_roofline(args.config, body, devb, flush, n, e, w["T"], net0.P, 1 if w["variant"] == "normed" else 0)
    others = None
    if rank == 0 and world == 1 and not args.no_large and not args.no_roofline:
        roof_large = roofline_large(dev, flush)
        if args.config == "qm9":
            others = other_config_points(dev, flush)
    # every rank replays the step here (it contains the gradient all-reduce when N > 1); rank 0 keeps the count
    launches = count_launches(run_resident)
    if args.timeline and world > 1 and rank == 0:
        # (the replayed step holds the fused all-reduce: one rank replaying alone would wait for its peers for ever)
        sys.stderr.write("bench.py: --timeline is a single-GPU option, skipped at %d GPUs\n" % world)
    if args.timeline and rank == 0 and world == 1:
        dump_timeline(run_resident, args.timeline)
        if gs_io is not None:       # the e2e step (transfers inside the graph)
            dump_timeline(gs_io.replay, args.timeline + ".e2e")
    if gs is not None:
        gs.check()
        if gs_io is not None:
            gs_io.check()

    if rank != 0:
        return
    gb_in = (sum(f.numel() for f in host_flat.values()) if use_graph
             else sum(v.numel() * v.element_size() for v in host.values()))
    line = {
        "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "graphs_per_gpu": B, "atoms_per_gpu": n, "directed_edges_per_gpu": e,
                   "optimizer": "Adam", "loss": "MSE", "cuda_graph": bool(use_graph), "l2_flush": "256 MB write between timed iterations",
                   "parallelism": "dp%d" % world,
                   "input_prep": ("compaction + de-duplication + type sort of batch k+1 run as a parallel branch of batch "
                                  "k's captured step (GraphedStep(pipeline_prep=True)): one prep and one train step per "
                                  "replay" if (gs is not None and gs.pipeline_prep) else "at the head of every step"),
                   "grad_allreduce": ("fused into the Adam launch over NVLink peer memory (k_adam_ddp)" if fused_ddp else
                                      ("NCCL, one flat bucket" if world > 1 else "none (1 GPU)"))},
        "e2e": {"value": world * B * args.steps / (ms2 * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": gb_in,
                "d2h_bytes_per_step": 4, "ms_per_step": ms2 / args.steps,
                "pipeline": ("one CUDA graph per step (GraphedStep(host_io=True)): device copy staging -> graph inputs, the step, "
                             "and the H2D of the host's pinned batch into the staging area as a parallel branch of the same "
                             "graph; every step's loss copied D2H (pinned) inside the timed region and read by the host one step "
                             "later" if (use_graph and gs_io is not None) else
                             "H2D of the next padded batch on a copy stream (pinned -> staging) overlapped with the current "
                             "step; D2D staging -> graph inputs; every step's loss copied D2H (pinned) inside the step and read by the host "
                             "one step later") if use_graph else "serial H2D; loss copied D2H every step, read one step later"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "median_ms_per_step": med_ms, "median_value": world * B / (med_ms * 1e-3),
    }
    if roof_large is not None:
        line["roofline_large"] = roof_large
    if others is not None:
        line["other_configs"] = others
    if ms3 is not None:
        line["e2e_device_collate"] = {
            "value": world * B * args.steps / (ms3 * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": rb_bytes,
            "d2h_bytes_per_step": 4, "ms_per_step": ms3 / args.steps,
            "note": "same as e2e, but the host ships the batch ragged (mpnn_b200.loader.RaggedBatch) and "
                    "mpnn_collate_ragged writes the reference's padded tensors on the GPU"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.config, budget_s=20.0)
    print(json.dumps(line))


def _graph_time(fn, flush, s_, reps=25):
    """median CUDA-event time (ms) of `fn()` replayed from its own CUDA graph (no Python / launch gaps inside the timed
    region), L2 flushed before every replay.  `s_`: the (non-default) stream the op's forward ran on -- autograd runs a
    node's backward on its forward's stream, so a captured backward must be captured on that stream."""
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s_)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s_):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    tt = []
    for _ in range(reps):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        tt.append(a.elapsed_time(b))
    return float(np.median(tt))


def survey_step_work(n, e, d, P, n_bn):
    """SURVEY.md 8d, forward, per message-passing step: FLOPs F_step and HBM bytes Q_step of the reference formulation
    (n real atoms, e directed bonds); fwd+bwd = 3 F_step and 2.5 Q_step."""
    F = 2.0 * e * P * d + 2.0 * n * P * d * d + 14.0 * n * d * d + 30.0 * n * d
    Q = 4.0 * e * P + 8.0 * e + 4.0 * n + 12.0 * n * d + 4.0 * (P * d * d + 6 * d * d + 8 * d) + 16.0 * n * d * n_bn
    return F, Q


def _ncu_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the MP-step kernels, from the committed `ncu --set full`
    summary of this workload (profiles/r02_mp_step_traffic.json, written by tools/ncu_traffic.py); None if not captured"""
    p = os.path.join(ROOT, "profiles", "r02_mp_step_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(tag)
    except Exception:
        return None


def mp_step_roofline(tag, body, devb, flush, n, e, T, P, n_bn):
    """SURVEY 8d's per-message-passing-step figure, measured live: the step kernels alone (message function +
    aggregation + GRU update [+ masked batch norm], forward and backward incl. the table gradients; NOT the
    compaction, the edge networks on the distinct rows or the readout, which SURVEY 8d reports per forward), replayed
    from their own CUDA graphs with the L2 flushed, against the MEASURED peaks.
    achieved = max(3 F_step / peak_flops, 2.5 Q_step / peak_bw) / t_step(fwd + bwd), as a bandwidth."""
    from mpnn_b200 import functional as Fn, graph, modules as M
    peaks = load_peaks()
    afm, bfm, adj, mask = devb["afm"], devb["bfm"], devb["adj"], devb["mask"]
    B, N, d = afm.shape
    nets = list(body.mfs) if hasattr(body, "mfs") else [body.mf]
    shared = not hasattr(body, "mfs")
    graph.clear_cache()
    el = graph.edge_list_for(bfm, adj)
    if el.typed().type_ptr is None:     # non-categorical bond rows: the per-edge path, no typed step kernels to time
        return None
    with torch.no_grad():
        tabs = [net._compute_table(el) for net in nets]
    tables = [t[0].detach().requires_grad_(True) for t in tabs]
    tablesT = [t[1].detach() for t in tabs]
    cell = body.uf.gru_cell
    ws = [p.detach().clone().requires_grad_(True) for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)]
    h0 = afm.reshape(-1, d)
    m1 = mask.reshape(-1)
    steps = T if not shared else 1      # shared edge network without chained state (autoencoder): one effective step
    prev = Fn.SIDE_STREAM_ENABLED
    Fn.SIDE_STREAM_ENABLED = False      # everything on the timed stream
    try:
        if Fn.chain_supported(d, steps) and d <= 32:
            bn = [dict(kind=1 if n_bn else 0, training=1, eps=1e-6, momentum=0.0, affine=False) for _ in range(steps)]
            tl = [tables[t % len(tables)] for t in range(steps)]
            tlT = [tablesT[t % len(tables)] for t in range(steps)]
            kernels = "k_chain_fwd / k_chain_bwd (+ k_tmsg_bwd_table, reduce): the whole T-step loop per launch"

            def fwd():
                return Fn.ChainFn.apply(h0, h0, m1, el, bn, *(ws + tl + tlT))
        else:
            kernels = ("k_tc_edge_gemm -> k_tc_gru_fwd (aggregation + GRU) | k_tc_gru_param_point + k_tc_gru_data_grad (d <= 64; "
                       "k_gru_point_bwd5 + grouped products above), k_tc_table_grad")

            # as modules._wide_chain runs it: on the real rows only (the gather / scatter of the node tensors happens
            # once per forward pass, outside the step)
            elc, _real, hc, mc = M.compact_nodes(afm, mask, el)

            def fwd():
                Mm = Fn.TypedMessageTCFn.apply(hc, tables[0], tablesT[0], elc, True, d, d)
                return Fn.GRUFn.apply(Mm, hc, mc, ws[0], ws[1], ws[2], ws[3], None)
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            out = fwd()
            gout = torch.randn_like(out)
        torch.cuda.synchronize()
        leaves = ws + tables[:max(1, min(len(tables), steps))]

        def bwd():
            torch.autograd.grad(out, leaves, gout, retain_graph=True, allow_unused=True)

        ms_f = _graph_time(lambda: fwd(), flush, s_)
        ms_b = _graph_time(bwd, flush, s_)
    finally:
        Fn.SIDE_STREAM_ENABLED = prev
        graph.clear_cache()
    F, Q = survey_step_work(n, e, d, P, n_bn)
    t_step = (ms_f + ms_b) * 1e-3 / steps
    t_hbm = 2.5 * Q / (peaks["hbm_gbs"] * 1e9)
    # The step kernels are HBM / latency-bound in this implementation: with the bond rows de-duplicated the contraction is
    # 2 e d^2 FLOPs per step (arithmetic intensity d/4 FLOP/B), not SURVEY's 2 n P d^2 -- the tensor-pipe roofline of the
    # reference formulation is reported for completeness only (`frac_tensor_reference_formulation`, can exceed 1).
    t_tc = 3.0 * F / (peaks["bf16_tflops"] * 1e12)
    frac = t_hbm / t_step
    ours_bytes = (12.0 * e + 4.0 * n + 12.0 * n * d + 16.0 * n * d * n_bn) * 2.5   # typed formulation: uid instead of x_e rows
    extra = {}
    if frac > 1.0:
        extra["frac_exceeds_one"] = ("SURVEY 8d's Q_step counts 4eP bytes of per-edge trunk rows (P = %d here); with the bond rows "
                                     "de-duplicated this implementation never moves them, so the step finishes faster than that "
                                     "traffic could be streamed: read frac_hbm_typed_formulation for the bytes it does move" % P)
    return {**extra, "what": "one message-passing step, forward + backward (SURVEY 8d), " + tag, "kernels": kernels,
            "bound": "hbm", "achieved": 2.5 * Q / t_step / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": frac,
            "traffic": _ncu_traffic(tag), "peak_source": peaks["source"] + " (copy bandwidth)",
            "ms_per_step_fwd": ms_f / steps, "ms_per_step_bwd": ms_b / steps, "steps_per_launch": steps,
            "algorithmic_bytes_per_step_fwd_bwd": 2.5 * Q, "algorithmic_flops_per_step_fwd_bwd": 3.0 * F,
            "frac_tensor_reference_formulation": t_tc / t_step,
            "typed_formulation_bytes_per_step_fwd_bwd": ours_bytes,
            "frac_hbm_typed_formulation": ours_bytes / (peaks["hbm_gbs"] * 1e9) / t_step,
            "atoms": n, "directed_edges": e, "d": d, "P": P,
            "note": "algorithmic work = SURVEY 8d's per-step formulas on the real atoms / bonds (reference formulation: "
                    "4eP bytes of trunk rows, 2nPd^2 contraction FLOPs; this implementation reads 4e bytes of type ids "
                    "and does 2ed^2 FLOPs instead, `typed_formulation_*`)"}


def roofline_large(dev, flush, points=((16384, 64), (16384, 128), (16384, 256))):
    """the same per-step figure where HBM is the limit: basic_graph_autoencoder (BASELINE configs[4]) at B = 16 384"""
    from mpnn_b200 import graph, synthetic
    out = []
    for Bn, h in points:
        try:
            w = dict(WORKLOADS["autoenc"], B=Bn, d=h, out=2 * h, targets=2 * h)
            batch = synthetic.make_batch("autoenc", B=Bn, d=h)
            devb = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask")}
            body, _ = build_model(w, dev)
            net = body.mf
            r = mp_step_roofline("autoenc_B%d_d%d" % (Bn, h), body, devb, flush, batch["n_atoms"], batch["n_edges"], 1,
                                 net.P, 0)
            r["graphs"] = Bn
            out.append(r)
            del devb, body, batch
            graph.clear_cache()
            torch.cuda.empty_cache()
        except Exception as ex:   # a point that does not fit is reported, not hidden
            out.append({"what": "autoenc_B%d_d%d" % (Bn, h), "error": repr(ex)[:300]})
    return out


def other_config_points(dev, flush, names=("zinc", "affinity_ecfp", "lipo"), steps=20):
    """the train step of BASELINE.json's other single-GPU configs (configs[0], [2], [3]) on this box, same method as the
    headline (captured step, L2 flushed between iterations, CUDA events, median): parity-tested workloads, not the
    metric line -- reported so that every config of the baseline has a driver-run number"""
    from mpnn_b200 import graph
    from mpnn_b200.graphs import GraphedStep
    from mpnn_b200.optim import FusedAdam
    out = []
    for name in names:
        try:
            w = dict(WORKLOADS[name])
            batch = make_workload_batch(name, w, 0)
            devb = {k: torch.from_numpy(batch[k]).to(dev) for k in ("afm", "bfm", "adj", "mask", "labels")}
            body, head = build_model(w, dev)
            opt = FusedAdam(list(body.parameters()) + list(head.parameters()), lr=1e-3)

            def step(b):
                graph.clear_cache()
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.mse_loss(head(body(b["afm"], b["bfm"], b["adj"], b["mask"])), b["labels"])
                loss.backward()
                opt.step()
                return loss

            use_graph = w.get("graph", True)
            gs = GraphedStep(step, devb, warmup=3) if use_graph else None
            run = gs.replay if gs is not None else (lambda: step(devb))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for i in range(steps):
                flush.fill_(i & 1)
                ev[i][0].record()
                run()
                ev[i][1].record()
            torch.cuda.synchronize()
            if gs is not None:
                gs.check()
            t = float(np.median([a.elapsed_time(b) for a, b in ev]))
            out.append({"workload": w["desc"], "config": name, "graphs": w["B"], "median_ms_per_step": t,
                        "graphs_per_s": w["B"] / (t * 1e-3), "cuda_graph": bool(use_graph), "steps": steps,
                        "launches": count_launches(run)})
            del gs, body, head, opt, devb
            graph.clear_cache()
            torch.cuda.empty_cache()
        except Exception as ex:
            out.append({"config": name, "error": repr(ex)[:300]})
    return out


def dump_timeline(fn, path):
    """warm per-kernel timeline of one step (CUPTI through torch.profiler): start offset, duration, stream, name"""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    with open(path, "w") as fh:
        fh.write("# start_us dur_us end_us stream name\n")
        for e in evs:
            a, b = e.time_range.start - t0, e.time_range.end - t0
            fh.write("%8.1f %7.1f %8.1f %4s %s\n" % (a, b - a, b, getattr(e, "device_resource_id", "?"),
                                                     e.name.replace("(anonymous namespace)::", "")[:90]))


def count_launches(fn):
    """number of kernels launched by one step (torch profiler, CUDA activity)"""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        return int(sum(ev.count for ev in prof.key_averages()
                       if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.key.lower()
                       and "memset" not in ev.key.lower()))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): same model, same synthetic batch, torch CPU fp32, all host cores
# ---------------------------------------------------------------------------------------------------
def _oracle_step_fn(config, B):
    from mpnn_b200 import synthetic
    from oracle import mpnn_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import leaf_sd
    w = WORKLOADS[config]
    batch = synthetic.make_batch(w.get("synth", config), B=B)
    body, head = build_model(w, torch.device("cpu"))
    sd = leaf_sd({k: v.detach().clone() for k, v in body.state_dict().items()})
    leaves, seen = [], set()
    for v in list(sd.values()) + list(head.parameters()):
        if v.dtype.is_floating_point and v.requires_grad and id(v) not in seen:
            seen.add(id(v))
            leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-3)
    t = {k: torch.from_numpy(batch[k]) for k in ("afm", "bfm", "adj", "mask")}
    labels = torch.from_numpy(batch["labels"]) if config in ("qm9", "qm9_continuous") else (
        torch.zeros(B, t["afm"].shape[1], w["targets"]) if w.get("per_atom") else torch.zeros(B, w["targets"]))
    buffers = {}

    def step():
        opt.zero_grad(set_to_none=True)
        a = (t["afm"], t["bfm"], t["adj"], t["mask"])
        if w["variant"] == "normed":
            y = O.normed_basic_model(*a, sd=sd, steps=w["T"])
        elif w["variant"] == "lipo":
            y = O.lipo_model(*a, sd=sd, steps=w["T"], buffers=buffers)
        elif w["variant"] == "att":
            y = O.att_model(*a, sd=sd, steps=w["T"], s2v_steps=100, agg="adj")
        elif w["variant"] == "normed_encoded":
            y = O.normed_encoded_model(*a, sd=sd, steps=w["T"], buffers=buffers)
        elif w["variant"] == "normed_encoded_ecfp":
            y = O.normed_encoded_ecfp_model(*a, sd=sd, steps=w["T"], buffers=buffers)
        else:
            y = O.basic_model(*a, sd=sd, steps=w["T"], chain_state=False)
        loss = torch.nn.functional.mse_loss(head(y), labels)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step, B


def cpu_baseline(config, budget_s=20.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = 32 if config != "lipo" else 32
    step, B = _oracle_step_fn(config, B)
    step()  # warm-up
    t0 = time.perf_counter()
    k = 0
    while True:
        step()
        k += 1
        if time.perf_counter() - t0 > budget_s or k >= 50:
            break
    dt = time.perf_counter() - t0
    return {"value": B * k / dt, "unit": "graphs/s", "cores": cores, "kind": "port",
            "sample": "%d fwd+bwd+Adam steps of the oracle port (dense B*N*N edge embedding, as the reference) on a "
                      "B=%d slice of the workload, torch CPU fp32, %d threads" % (k, B, cores),
            "ms_per_step": dt / k * 1e3,
            "config": {"same_config": False, "graphs_per_step": B, "warmup_steps": 1,
                       "note": "graphs/s normalises the batch; the GPU arm runs the workload's full batch"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOADS[args.config]
    B = 32
    step, B = _oracle_step_fn(args.config, B)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = B * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "graphs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": "B=%d graphs per step (bounded CPU sample of the same workload)" % B,
                   "same_config": False, "graphs_per_step": B, "warmup_steps": max(1, min(args.warmup, 2))},
        "cpu_baseline": {"value": v, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": "%d steps at B=%d, oracle port of the reference's PyTorch CPU path" % (steps, B)},
        "e2e": {"value": v, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--no-roofline", action="store_true", help="skip the roofline measurements (profiling runs)")
    ap.add_argument("--no-large", action="store_true", help="skip the roofline_large points (autoenc B=16384, d=64/256)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="qm9", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stock-adam", action="store_true", help="torch.optim.Adam(fused) instead of mpnn_b200.optim.FusedAdam")
    ap.add_argument("--stock-head", action="store_true", help="keep the head + loss as stock torch modules")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N>1: NCCL all-reduce + Adam instead of the fused kernel")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one captured CUDA graph")
    ap.add_argument("--no-host-io", action="store_true",
                    help="e2e: separate H2D / D2D launches around the step's graph instead of transfers inside the graph")
    ap.add_argument("--no-pipeline-prep", action="store_true",
                    help="compaction / de-duplication at the head of each step instead of one step ahead")
    ap.add_argument("--batch", type=int, default=0, help="graphs per GPU (default: the workload's BASELINE.json batch)")
    ap.add_argument("--timeline", default="", help="write a warm per-kernel timeline of one step to this file")
    ap.add_argument("--hidden", type=int, default=0, help="feature width d (autoenc sweep of BASELINE configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
