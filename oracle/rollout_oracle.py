"""CPU restatement of the reference's attention rollout -> CLS-row token score  --  TEST / BENCH INFRASTRUCTURE ONLY.

SURVEY.md §8(f) next #1: the producer of ``cls_token_attn`` (the score the prototype head's selection consumes).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this module; the product
path (protopformer_b200/) never does.

Reference code followed (file:line relative to /root/reference):
  * tools/deit_models_attn.py:99-124   ``MyVisionTransformer.attn_rollout``: for every layer, head fusion (mean over
    heads, :102-104), discard of the ``int(T*T*discard_ratio)`` smallest entries of the flattened fused map
    (``topk(..., largest=False)`` + ``scatter_(.., 0)``, :109-113), ``a = (A + 0.2 I) / 1.2`` (:118-119), row
    normalisation (:121), ``result = a @ result`` (:123).
  * tools/deit_models_attn.py:222-226  the consumer: ``cls_token_attn = attn_rollout[:, 0, 1:]`` -- ONLY row 0 of the
    (T,T) product is used, so the chain of L batched T^3 matmuls collapses to L vector-matrix products
    ``v <- v @ a_l`` walked from the last layer to the first (``rollout_cls_row`` below).
  * tools/cait_models_attn.py:223-261  CaiT variant: same per-layer processing on (T-1)x(T-1) patch layers, the start
    vector is the mean of the processed class-attention rows instead of e_0 (``v0`` argument).

Pin: tests/golden/rollout_*.npz hold the outputs of the reference's own ``attn_rollout`` (imported unmodified through
oracle/ref_harness.py, torch 2.11 CPU fp32) on seeded inputs; tests/test_rollout_oracle.py holds this restatement to
them.  Tie rule (the reference leaves it to ATen's topk, i.e. unspecified): among entries EQUAL to the k-th smallest
value the ones with the lowest flat index are discarded first; fixtures are checked to be tie-free at the threshold.
"""
from __future__ import annotations

import torch


def synth_attention(L: int, B: int, H: int, T: int, seed: int = 1, sharp: float = 2.0):
    """L attention tensors (B,H,T,T): row-softmax of scaled Gaussian logits (rows sum to 1, peaky like real maps)."""
    g = torch.Generator().manual_seed(4000 + seed)
    return [torch.softmax(sharp * torch.randn(B, H, T, T, generator=g), dim=-1) for _ in range(L)]


def fuse_heads(attn: torch.Tensor, head_fusion: str = "mean") -> torch.Tensor:
    """deit_models_attn.py:102-107."""
    if head_fusion == "mean":
        return attn.mean(dim=1)
    if head_fusion == "max":
        return attn.max(dim=1)[0]
    if head_fusion == "min":
        return attn.min(dim=1)[0]
    raise ValueError(head_fusion)


def discard_count(n_elems: int, discard_ratio: float) -> int:
    """deit_models_attn.py:110: ``int(flat.shape[-1] * discard_ratio)`` (Python float arithmetic, truncation)."""
    return int(n_elems * discard_ratio)


def discard_smallest(fused: torch.Tensor, k: int) -> torch.Tensor:
    """Zero the k smallest entries of each image's flattened map (deit_models_attn.py:109-113); ties at the
    threshold: lowest flat index first (stable sort)."""
    B = fused.shape[0]
    flat = fused.reshape(B, -1).clone()
    if k > 0:
        order = torch.sort(flat, dim=-1, stable=True)[1][:, :k]
        flat.scatter_(1, order, 0.0)
    return flat.reshape(fused.shape)


def process_layer(attn: torch.Tensor, discard_ratio: float = 0.9, head_fusion: str = "mean",
                  identity_w: float = 0.2) -> torch.Tensor:
    """(B,H,R,T) attention -> row-stochastic (B,R,T) matrix a_l (deit_models_attn.py:102-121; the CaiT variant's
    ``I[:R]`` for non-square class-attention maps, cait_models_attn.py:239-240)."""
    fused = fuse_heads(attn, head_fusion)
    R, T = fused.shape[-2:]
    kept = discard_smallest(fused, discard_count(R * T, discard_ratio))
    eye = torch.eye(T, dtype=fused.dtype)[:R]
    a = (kept + identity_w * eye) / (1.0 + identity_w)
    return a / a.sum(dim=-1).unsqueeze(dim=-1)


def rollout_full(all_attn, discard_ratio: float = 0.9, head_fusion: str = "mean") -> torch.Tensor:
    """The reference's algorithm as written: full (B,T,T) product chain (deit_models_attn.py:100, 123)."""
    B, T = all_attn[0].shape[0], all_attn[0].shape[-1]
    result = torch.eye(T, dtype=all_attn[0].dtype).unsqueeze(0).repeat(B, 1, 1)
    for attn in all_attn:
        result = torch.matmul(process_layer(attn, discard_ratio, head_fusion), result)
    return result


def rollout_cls_row(all_attn, discard_ratio: float = 0.9, head_fusion: str = "mean", v0: torch.Tensor | None = None,
                    drop_first: bool = True, dtype=torch.float32) -> torch.Tensor:
    """Row 0 of ``rollout_full`` without forming the product: v = e_0 (or ``v0`` (B,T)); for l = L-1 .. 0:
    v <- v @ a_l; returns v[:, 1:] (= ``attn_rollout[:, 0, 1:]``, deit_models_attn.py:226) or all of v."""
    B, T = all_attn[0].shape[0], all_attn[0].shape[-1]
    if v0 is None:
        v = torch.zeros(B, T, dtype=dtype)
        v[:, 0] = 1.0
    else:
        v = v0.to(dtype).clone()
    for attn in reversed(list(all_attn)):
        a = process_layer(attn, discard_ratio, head_fusion).to(dtype)
        v = torch.einsum("bi,bij->bj", v, a)
    return v[:, 1:] if drop_first else v


def threshold_tie_free(all_attn, discard_ratio: float = 0.9, head_fusion: str = "mean") -> bool:
    """True when, in every (layer, image), the k-th and (k+1)-th smallest fused values differ (so every topk
    implementation discards the same set)."""
    for attn in all_attn:
        fused = fuse_heads(attn, head_fusion)
        B = fused.shape[0]
        flat = fused.reshape(B, -1)
        k = discard_count(flat.shape[-1], discard_ratio)
        if k <= 0 or k >= flat.shape[-1]:
            continue
        s = torch.sort(flat, dim=-1)[0]
        if bool((s[:, k - 1] == s[:, k]).any()):
            return False
    return True


def rollout_cait(all_attn, pre_layer_num: int, discard_ratio: float = 0.9, head_fusion: str = "mean") -> torch.Tensor:
    """tools/cait_models_attn.py:223-261 as written: every map is processed (:225-246), the first ``pre_layer_num``
    form the (T,T) product (:251-253), the mean of the remaining class-attention rows without the CLS column (:255-257)
    multiplies it (:258); returns ``cls_result[:, 0]`` (B,T) -- what :328-330 feed the token selection."""
    proc = [process_layer(a, discard_ratio, head_fusion) for a in all_attn]
    patch, cls = proc[:pre_layer_num], proc[pre_layer_num:]
    B, T = patch[0].shape[0], patch[0].shape[-1]
    result = torch.eye(T, dtype=patch[0].dtype).unsqueeze(0).repeat(B, 1, 1)
    for a in patch:
        result = torch.matmul(a, result)
    cls_result = torch.cat(cls, dim=1).mean(dim=1, keepdim=True)[:, :, 1:]
    return (cls_result @ result)[:, 0]


def synth_cait_attention(n_patch: int, n_cls: int, B: int, H: int, T: int, seed: int = 1, sharp: float = 2.0):
    """n_patch maps (B,H,T,T) followed by n_cls class-attention maps (B,H,1,T+1)."""
    g = torch.Generator().manual_seed(5000 + seed)
    maps = [torch.softmax(sharp * torch.randn(B, H, T, T, generator=g), dim=-1) for _ in range(n_patch)]
    maps += [torch.softmax(sharp * torch.randn(B, H, 1, T + 1, generator=g), dim=-1) for _ in range(n_cls)]
    return maps
