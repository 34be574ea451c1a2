"""GPU suite: pph_adamw_step (csrc/pph_adamw.cu, protopformer_b200/optim.py) against torch.optim.AdamW and the oracle."""
import pytest
import torch

from oracle import adamw_oracle as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LRS = {"add_on_layers": 3e-3, "prototype_vectors": 3e-3}


@pytest.mark.parametrize("shapes", [
    {"Wa": (24, 16), "ba": (24,), "P": (40, 24), "Pg": (20, 24)},
    {"Wa": (192, 192), "ba": (192,), "P": (2000, 192), "Pg": (2000, 192)},          # CUB shape: 805 k elements
    {"Wa": (5, 3), "ba": (5,), "P": (2049, 1), "Pg": (1, 7)},                        # ragged tails, 1-block tensors
])
def test_fused_adamw_matches_torch_adamw(shapes):
    from protopformer_b200.optim import FusedHeadAdamW
    t, groups = A.head_groups_like_reference(shapes, LRS, 0.05)
    ref_opt = torch.optim.AdamW(groups, weight_decay=0.05, eps=1e-8)
    for p in t.values():
        p.requires_grad_(True)
    d = {k: v.detach().clone().to(DEV) for k, v in t.items()}
    dgroups = [{"params": [d["Wa"], d["ba"]], "lr": 3e-3, "weight_decay": 1e-3},
               {"params": [d["P"]], "lr": 3e-3, "weight_decay": 0.05},
               {"params": [d["Pg"]], "lr": 3e-3, "weight_decay": 0.05}]
    opt = FusedHeadAdamW(dgroups, weight_decay=0.05, eps=1e-8)
    g = torch.Generator().manual_seed(11)
    for step in range(1, 7):
        if step == 3:
            for gr in ref_opt.param_groups:
                gr["lr"] = 5e-4
            for gr in opt.param_groups:
                gr["lr"] = 5e-4
            opt.sync_hyper()
        for k, p in t.items():
            p.grad = 0.1 * torch.randn(p.shape, generator=g)
            d[k].grad = p.grad.to(DEV)
        ref_opt.step()
        opt.step()
        assert opt.step_count == step
        for k in t:
            # fp32 elementwise update: same operations, possibly different FMA contraction -> 1e-6 relative
            assert torch.allclose(d[k].cpu(), t[k].detach(), rtol=1e-6, atol=1e-7), (k, step)
    sd, rsd = opt.state_dict(), ref_opt.state_dict()
    assert [g["params"] for g in sd["param_groups"]] == [g["params"] for g in rsd["param_groups"]]
    for i in rsd["state"]:
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"])
        assert torch.allclose(sd["state"][i]["exp_avg_sq"].cpu(), rsd["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_fused_adamw_in_cuda_graph_follows_lr_and_reads_flat_gradients():
    """The launch is replayed from a CUDA graph: step count and learning rate come from device memory."""
    from protopformer_b200.optim import FusedHeadAdamW
    g = torch.Generator().manual_seed(3)
    p_host = torch.rand(1000, 64, generator=g)
    grad_host = 0.05 * torch.randn(1000, 64, generator=g)
    p = p_host.clone().to(DEV)
    flat = grad_host.clone().to(DEV)                    # stands for a view of the all-reduce buffer
    opt = FusedHeadAdamW([p], lr=3e-3, weight_decay=0.05, grads=[flat])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()                                       # warm-up launch outside capture = update 1
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    pr = p_host.clone()
    m, v = torch.zeros_like(pr), torch.zeros_like(pr)
    A.adamw_step(pr, grad_host, m, v, 1, 3e-3, 0.05)
    for t in range(2, 6):
        lr = 3e-3 if t < 4 else 1e-3
        if t == 4:
            opt.param_groups[0]["lr"] = 1e-3
            opt.sync_hyper()
        graph.replay()
        A.adamw_step(pr, grad_host, m, v, t, lr, 0.05)
    torch.cuda.synchronize()
    assert opt.step_count == 5
    assert torch.allclose(p.cpu(), pr, rtol=1e-6, atol=1e-7)


def test_fused_adamw_refuses_cpu_tensors():
    from protopformer_b200.optim import FusedHeadAdamW
    with pytest.raises(ValueError):
        FusedHeadAdamW([torch.zeros(4, 4)])


def test_zero_grad_keeps_the_flat_all_reduce_views_attached():
    """optimizer.zero_grad() (set_to_none by default) must not detach the gradients from the flat
    all-reduce buffer the optimizer and the exchange read."""
    from protopformer_b200.dist import FlatGradReducer
    from protopformer_b200.optim import FusedHeadAdamW
    dev = torch.device("cuda:0")
    ps = [torch.randn(7, 5, device=dev, requires_grad=True), torch.randn(11, device=dev, requires_grad=True)]
    red = FlatGradReducer([("a", ps[0]), ("b", ps[1])])
    opt = FusedHeadAdamW([{"params": ps, "lr": 1e-3, "weight_decay": 0.0}], grads=red.views)
    (ps[0].sum() * 2.0 + ps[1].sum()).backward()
    assert float(red.flat.abs().sum()) > 0
    opt.zero_grad()
    assert float(red.flat.abs().sum()) == 0.0
    red.check_attached()                                    # still the same storage
    (ps[0].sum() * 3.0).backward()
    assert torch.equal(red.views[0], torch.full_like(ps[0], 3.0))
    ps[0].grad = None                                       # a caller that dropped it anyway ...
    with pytest.raises(RuntimeError):
        red.check_attached()                                # ... is told so instead of averaging a stale buffer
    opt.zero_grad()
    red.check_attached()                                    # and zero_grad() re-attaches
