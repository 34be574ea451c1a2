"""Synthetic inputs for the prototype-head path (TEST / BENCH INFRASTRUCTURE, not product code).

Every tensor is drawn from an explicit ``torch.Generator`` so the same (shape, seed) pair yields the same
bytes in this container (where the golden fixtures are produced from the real reference) and on the GPU box
(where only the CPU restatement and the CUDA path run).  Distributions follow SURVEY.md §8(d):

* tokens  (B, 1+N, Din)  ~ N(0,1)              -- LayerNorm-like backbone output (protopformer.py:155)
* scores  (B, N)         per-image permutation  -- exactly tie-free CLS-attention rollout (protopformer.py:157)
* labels  (B,)           uniform ints in [0,C)
* prototype_vectors / prototype_vectors_global ~ U[0,1)   (protopformer.py:115-119)
* add-on 1x1 conv weight kaiming-normal fan_out, bias small  (protopformer.py:388-395; bias is 0 at init in the
  reference, a small non-zero bias is used here so the bias path is exercised)
* last layers +1 on the own class / -0.5 elsewhere         (protopformer.py:367-386)

``proto_mode="matched"`` builds the "trained-like" distribution: prototypes are copies of sigmoided add-on
outputs of random tokens plus N(0, sigma) noise, so some distances are small (cancellation regime).
"""
from __future__ import annotations

import dataclasses
import math

import torch


@dataclasses.dataclass(frozen=True)
class HeadShape:
    """Shape of one prototype-head problem (names follow SURVEY.md §8)."""

    name: str
    B: int          # images
    N: int          # patch tokens (perfect square, 196 for 224x224 / patch 16)
    Din: int        # backbone width
    D: int          # prototype dim
    K: int          # reserve_token_nums[-1] (perfect square)
    P: int          # local prototypes
    Pg: int         # global prototypes = C * global_proto_per_class
    C: int          # classes
    global_coe: float = 0.5
    ppc_cov_thresh: float = 1.0
    ppc_mean_thresh: float = 2.0

    @property
    def m(self) -> int:
        return self.P // self.C

    def with_batch(self, B: int) -> "HeadShape":
        return dataclasses.replace(self, B=B)


# BASELINE.json configs (head-only view); B is the per-GPU batch of that config.
SHAPES = {
    "cub_b8": HeadShape("cub_b8", 8, 196, 192, 192, 81, 2000, 2000, 200, 0.5, 1.0, 2.0),
    "cub_b64": HeadShape("cub_b64", 64, 196, 192, 192, 81, 2000, 2000, 200, 0.5, 1.0, 2.0),
    "dogs_b256": HeadShape("dogs_b256", 256, 196, 384, 384, 81, 1200, 600, 120, 0.5, 1.0, 2.0),
    "cars_b64": HeadShape("cars_b64", 64, 196, 192, 192, 121, 1960, 980, 196, 0.5, 1.0, 2.0),
    # corners of the BASELINE config 5 sweep (tokens 49-196, dim 192/384) at fixture-sized batches
    "sweep_k49": HeadShape("sweep_k49", 3, 196, 192, 192, 49, 1000, 500, 100, 0.5, 1.0, 2.0),
    "sweep_k196": HeadShape("sweep_k196", 2, 196, 192, 192, 196, 1000, 1000, 100, 0.5, 1.0, 2.0),
    "sweep_k144_d384": HeadShape("sweep_k144_d384", 2, 196, 384, 384, 144, 1000, 500, 100, 0.5, 1.0, 2.0),
    # small shapes the pure-python/float64 checks finish instantly on
    "tiny": HeadShape("tiny", 3, 16, 24, 16, 9, 20, 8, 4, 0.3, 0.2, 1.5),
    "small": HeadShape("small", 5, 49, 40, 32, 25, 60, 30, 6, 0.5, 0.5, 2.0),
}


def make_case(shape: HeadShape, seed: int = 1, proto_mode: str = "init", sigma: float = 0.05,
              heads: int = 0, dtype=torch.float32) -> dict:
    """Returns a dict of CPU tensors: tokens, scores, labels, P, Pg, Wa, ba, Wl, Wg (+ scores_h if heads>0)."""
    s = shape
    g = torch.Generator().manual_seed(1000 + seed)
    tokens = torch.randn(s.B, 1 + s.N, s.Din, generator=g)
    # tie-free scores: a permutation of 1..N per image, scaled to look like a probability row
    scores = torch.stack([(torch.randperm(s.N, generator=g) + 1).float() for _ in range(s.B)])
    scores = scores / float(s.N * (s.N + 1) // 2)
    labels = torch.randint(0, s.C, (s.B,), generator=g)
    gp = torch.Generator().manual_seed(2000 + seed)
    P = torch.rand(s.P, s.D, generator=gp)
    Pg = torch.rand(s.Pg, s.D, generator=gp)
    Wa = torch.randn(s.D, s.Din, generator=gp) * math.sqrt(2.0 / s.D)   # kaiming-normal, fan_out = D*1*1
    ba = 0.1 * torch.randn(s.D, generator=gp)
    m, mg = s.P // s.C, s.Pg // s.C
    Wl = torch.full((s.C, s.P), -0.5)
    Wg = torch.full((s.C, s.Pg), -0.5)
    for c in range(s.C):
        Wl[c, c * m:(c + 1) * m] = 1.0
        Wg[c, c * mg:(c + 1) * mg] = 1.0
    if proto_mode == "matched":
        # prototypes = sigmoided add-on features of random tokens (+ noise): small distances appear
        z = torch.sigmoid(tokens.reshape(-1, s.Din) @ Wa.t() + ba)
        pick = torch.randint(0, z.shape[0], (s.P,), generator=gp)
        P = (z[pick] + sigma * torch.randn(s.P, s.D, generator=gp)).clamp(0.0, 1.0)
        pick = torch.randint(0, z.shape[0], (s.Pg,), generator=gp)
        Pg = (z[pick] + sigma * torch.randn(s.Pg, s.D, generator=gp)).clamp(0.0, 1.0)
    elif proto_mode != "init":
        raise ValueError(proto_mode)
    case = dict(tokens=tokens, scores=scores, labels=labels, P=P, Pg=Pg, Wa=Wa, ba=ba, Wl=Wl, Wg=Wg)
    if heads > 0:
        # (B,H,N) positive per-head scores whose mean over H is exactly tie-free: perturb around `scores`
        # with zero-sum noise across heads, small enough not to reorder (gap between ranks is 1/sumN).
        gap = 1.0 / float(s.N * (s.N + 1) // 2)
        noise = torch.rand(s.B, heads, s.N, generator=g)
        noise = (noise - noise.mean(dim=1, keepdim=True)) * gap * 0.5  # zero-sum over heads: mean stays tie-free
        case["scores_h"] = (scores[:, None, :] + noise).contiguous()
    if dtype != torch.float32:
        case = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in case.items()}
    return case


def checksum(t: torch.Tensor) -> float:
    """Order-independent-ish fingerprint used to detect RNG drift between the fixture maker and the tests."""
    t = t.double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0) + 1.0
    return float((t * w).sum())
