"""2..8-GPU check of the fused all-reduce + Adam step (torchrun --nproc-per-node N tools/ddp_check.py):
(a) FlatGradAllReduce (NCCL) + FusedAdam vs (b) FusedAdam.enable_ddp(): parameters after 4 steps agree to fp32 round-off
(the two sum in different orders), and with (b) every rank holds bit-identical parameters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mpnn_b200 import dist as D, graph, synthetic
from mpnn_b200.dropin import reference_model, kaiming_init
from mpnn_b200.optim import FusedAdam

rank, world = D.init_from_env()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
b = synthetic.make_batch("qm9", B=64, seed_offset=rank)
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("afm", "bfm", "adj", "mask", "labels")}


def run(mode):
    torch.manual_seed(317)
    body = reference_model("normed", 16, 7, 16, 1, 64, message_steps=3)
    body.apply(kaiming_init)
    head = torch.nn.Linear(64, 12)
    body, head = body.to(dev), head.to(dev)
    params = list(body.parameters()) + list(head.parameters())
    opt = FusedAdam(params, lr=1e-3)
    ar = D.FlatGradAllReduce(params)
    if mode == "fused":
        assert opt.enable_ddp()
    for _ in range(4):
        graph.clear_cache()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(head(body(t["afm"], t["bfm"], t["adj"], t["mask"])), t["labels"])
        loss.backward()
        if mode == "nccl":
            ar()
        opt.step()
    torch.cuda.synchronize()
    return torch.cat([p.detach().reshape(-1) for p in params])


a = run("nccl")
f = run("fused")
err = float((a - f).abs().max() / a.abs().max())
gathered = [torch.empty_like(f) for _ in range(world)]
dist.all_gather(gathered, f)
same = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print("ddp_check world=%d: fused vs nccl max rel diff %.3e, ranks bit-identical: %s" % (world, err, same))
assert err < 1e-5 and same
dist.barrier()
dist.destroy_process_group()
