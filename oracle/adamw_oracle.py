"""CPU restatement of the optimizer tail for the head's parameter groups  --  TEST INFRASTRUCTURE ONLY.

The reference steps ``torch.optim.AdamW`` (tools/create_optimizer.py:92, built over the groups of
``split_weights`` :31-39; stepped at tools/engine_proto.py:76-78 through timm's NativeScaler).  The arithmetic lives in
PyTorch itself (reference pins pytorch==1.8.1, README.md:56; this container has torch 2.11 -- the decoupled update is
unchanged): tests/test_adamw_oracle.py pins ``adamw_step`` below against ``torch.optim.AdamW`` on CPU, and
tests/test_adamw_gpu.py holds the CUDA kernel (csrc/pph_adamw.cu) to both.
"""
from __future__ import annotations

import math

import torch


def adamw_step(p, g, m, v, t: int, lr: float, wd: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """One decoupled-weight-decay Adam update, in place, in the order torch's single-tensor path uses; t = 1, 2, ..."""
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
    return p


def head_groups_like_reference(shapes: dict, lrs: dict, weight_decay: float, seed: int = 0):
    """Random stand-ins for the head's tensors grouped as tools/create_optimizer.py:31-39 does (add-on layers with
    weight decay 1e-3; prototypes with the optimizer-level decay)."""
    g = torch.Generator().manual_seed(7000 + seed)
    t = {k: torch.rand(*s, generator=g) for k, s in shapes.items()}
    groups = [
        {"params": [t["Wa"], t["ba"]], "lr": lrs["add_on_layers"], "weight_decay": 1e-3},
        {"params": [t["P"]], "lr": lrs["prototype_vectors"], "weight_decay": weight_decay},
        {"params": [t["Pg"]], "lr": lrs["prototype_vectors"], "weight_decay": weight_decay},
    ]
    return t, groups
