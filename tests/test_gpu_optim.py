"""mpnn_b200.optim.FusedAdam (csrc/optim.cu) against torch.optim.Adam on the same parameters and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_fused_adam_matches_torch_adam(wd):
    from mpnn_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    shapes = [(49, 49), (256, 49), (256,), (16, 48), (48,), (1,), (3, 5, 7), (70000,)] + [(7, 3)] * 45   # > 40 tensors
    pa = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.Adam(pa, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    ob = FusedAdam(pb, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    for it in range(5):
        g = torch.Generator(device="cpu").manual_seed(it)
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(dev)
            x.grad = gr.clone()
            y.grad = gr.clone()
        oa.step()
        ob.step()
        for x, y in zip(pa, pb):
            assert torch.allclose(x, y, rtol=1e-5, atol=2e-6), (it, x.shape, float((x - y).abs().max()))
    assert float(ob.param_groups[0]["_mpnn_state"]["step"]) == 5.0
    st = ob.state[pb[0]]
    assert torch.allclose(st["exp_avg"], oa.state[pa[0]]["exp_avg"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(st["exp_avg_sq"], oa.state[pa[0]]["exp_avg_sq"], rtol=1e-5, atol=1e-7)


def test_fused_adam_in_cuda_graph():
    from mpnn_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    w = torch.randn(64, 12, device=dev).requires_grad_(True)
    w_ref = w.detach().clone().requires_grad_(True)
    x = torch.randn(256, 64, device=dev)
    opt = FusedAdam([w], lr=1e-2)
    ref = torch.optim.Adam([w_ref], lr=1e-2)

    def step(wt, o):
        o.zero_grad(set_to_none=True)
        loss = (x @ wt).pow(2).mean()
        loss.backward()
        o.step()
        return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step(w, opt)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(w, opt)
    for _ in range(4):
        g.replay()
    for _ in range(6):        # 2 eager warm-up steps + 4 replays (the capture itself executes nothing)
        step(w_ref, ref)
    torch.cuda.synchronize()
    assert torch.allclose(w, w_ref, rtol=1e-4, atol=1e-6)


def test_fused_adam_skips_parameters_without_gradient():
    """(the step count is one device counter per group: a skipped parameter keeps its value but does not lag behind in
    bias correction the way torch's per-parameter counters do -- documented difference)"""
    from mpnn_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    a = torch.randn(5, device=dev).requires_grad_(True)
    b = torch.randn(5, device=dev).requires_grad_(True)
    a0, b0 = a.detach().clone(), b.detach().clone()
    opt = FusedAdam([a, b], lr=1e-2)
    a.grad = torch.ones_like(a)
    opt.step()
    assert torch.equal(b, b0) and not torch.equal(a, a0)


def test_fused_adam_rejects_cpu_parameters():
    from mpnn_b200.optim import FusedAdam
    p = torch.randn(3, requires_grad=True)
    p.grad = torch.randn(3)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()
