"""CPU suite: the AdamW restatement against torch.optim.AdamW (the reference's optimizer, create_optimizer.py:92)."""
import torch

from oracle import adamw_oracle as A

SHAPES = {"Wa": (24, 16), "ba": (24,), "P": (40, 24), "Pg": (20, 24)}
LRS = {"add_on_layers": 3e-3, "prototype_vectors": 3e-3}


def test_restatement_matches_torch_adamw_over_steps():
    t, groups = A.head_groups_like_reference(SHAPES, LRS, 0.05)
    mine = {k: v.clone() for k, v in t.items()}
    m = {k: torch.zeros_like(v) for k, v in t.items()}
    v2 = {k: torch.zeros_like(v) for k, v in t.items()}
    for p in t.values():
        p.requires_grad_(True)
    opt = torch.optim.AdamW(groups, weight_decay=0.05, eps=1e-8)
    g = torch.Generator().manual_seed(5)
    hyp = {"Wa": (3e-3, 1e-3), "ba": (3e-3, 1e-3), "P": (3e-3, 0.05), "Pg": (3e-3, 0.05)}
    for step in range(1, 8):
        if step == 4:                      # a scheduler changes the learning rate between steps
            for gr in opt.param_groups:
                gr["lr"] = 1e-3
            hyp = {k: (1e-3, w) for k, (_, w) in hyp.items()}
        for k, p in t.items():
            p.grad = 0.1 * torch.randn(p.shape, generator=g)
            A.adamw_step(mine[k], p.grad, m[k], v2[k], step, *hyp[k])
        opt.step()
        for k in t:
            assert torch.allclose(mine[k], t[k].detach(), rtol=1e-6, atol=1e-7), (k, step)
