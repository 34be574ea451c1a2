"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from protopformer_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (shape key, batch override, seed, proto_mode, activation)   (mirrors tests/golden/make_golden.py)
GOLDEN_CASES = {
    "tiny_s1": ("tiny", None, 1, "init", "log"),
    "tiny_s2_linear": ("tiny", None, 2, "init", "linear"),
    "small_s1": ("small", None, 1, "init", "log"),
    "small_s3_matched": ("small", None, 3, "matched", "log"),
    "cub_b8_s1": ("cub_b8", None, 1, "init", "log"),
    "cub_b8_s2_matched": ("cub_b8", None, 2, "matched", "log"),
    "cars_b4_s1": ("cars_b64", 4, 1, "init", "log"),
    "dogs_b4_s1": ("dogs_b256", 4, 1, "init", "log"),
    # corners of the BASELINE config 5 sweep (tokens 49 / 196 = N / 144 with D = 384)
    "sweep_k49_s1": ("sweep_k49", None, 1, "init", "log"),
    "sweep_k196_s1": ("sweep_k196", None, 1, "init", "log"),
    "sweep_k144_d384_s1": ("sweep_k144_d384", None, 1, "init", "log"),
}

EXTRA_GOLDEN_CASES = {}          # (kept for the fixture generator's interface: every fixture is a first-class case now)


def load_golden(name):
    key, b, seed, mode, fn = GOLDEN_CASES[name]
    shape = synth.SHAPES[key]
    if b is not None:
        shape = shape.with_batch(b)
    case = synth.make_case(shape, seed=seed, proto_mode=mode)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    for k in ("tokens", "scores", "P", "Pg", "Wa", "ba"):
        assert abs(synth.checksum(case[k]) - float(g["chk_" + k])) <= 1e-9 * max(1.0, abs(float(g["chk_" + k]))), \
            f"synthetic input {k} drifted from the fixture"
    return shape, case, g, fn


def rel_close(a, b, rtol=1e-4, atol=1e-6):
    """|a-b| <= rtol*max(|a|,|b|) + atol elementwise (SURVEY.md §8(d) parity threshold)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return bool(((a - b).abs() <= rtol * torch.maximum(a.abs(), b.abs()) + atol).all())


def max_rel(a, b, atol=1e-6):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float(((a - b).abs() / (torch.maximum(a.abs(), b.abs()) + atol)).max())


def norm_rel(a, b):
    """max |a-b| / max |b|: the metric for gradients (many entries are ~0, elementwise-relative is meaningless)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def argmax_mismatch_outside_near_ties(arg_a, arg_b, near_tie):
    """#positions where two argmax maps differ, not counting the fixture's near-tie (b,p) pairs."""
    diff = torch.as_tensor(arg_a).long() != torch.as_tensor(arg_b).long()
    for b, p in np.asarray(near_tie).reshape(-1, 2):
        diff[b, p] = False
    return int(diff.sum())


class FakeAttnBlock(torch.nn.Module):
    """Seeded stand-in for a transformer block with the reference's calling convention ``blk(x, policy) -> (x, attn)``
    (tools/deit_models_attn.py:76-82): multi-head softmax attention whose keys are masked by ``policy`` (pruned tokens
    only attend to themselves, as :29-43 does), plus a residual.  TEST INFRASTRUCTURE for the backbone-loop tests."""

    def __init__(self, dim: int, heads: int, seed: int):
        super().__init__()
        g = torch.Generator().manual_seed(9000 + seed)
        self.heads = heads
        self.wq = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)
        self.wk = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)
        self.wv = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)

    def forward(self, x, policy):
        B, T, C = x.shape
        H, hd = self.heads, C // self.heads
        q = (x @ self.wq).reshape(B, T, H, hd).transpose(1, 2)
        k = (x @ self.wk).reshape(B, T, H, hd).transpose(1, 2)
        v = (x @ self.wv).reshape(B, T, H, hd).transpose(1, 2)
        logits = (q @ k.transpose(-2, -1)) * (3.0 * hd ** -0.5)
        mask = policy.reshape(B, 1, 1, T)
        mask = mask + (1.0 - mask) * torch.eye(T, device=x.device).view(1, 1, T, T)
        e = torch.exp(logits - logits.max(dim=-1, keepdim=True)[0]) * mask
        attn = e / e.sum(dim=-1, keepdim=True)
        return x + 0.5 * (attn @ v).transpose(1, 2).reshape(B, T, C), attn


class FakeDeit(torch.nn.Module):
    """``blocks`` + ``norm``: what forward_feature_mask_train_direct touches (tools/deit_models_attn.py:205-241)."""

    def __init__(self, dim: int, heads: int, depth: int):
        super().__init__()
        self.blocks = torch.nn.ModuleList([FakeAttnBlock(dim, heads, i) for i in range(depth)])
        self.norm = torch.nn.LayerNorm(dim)


class FakeClassAttnBlock(torch.nn.Module):
    """Seeded stand-in for CaiT's token-only block (tools/cait_models_attn.py:161-186 calling convention):
    ``blk(x, cls_tokens, policy) -> (cls_tokens, attn (B,H,1,1+N))``; pruned tokens (policy 0) get no attention."""

    def __init__(self, dim: int, heads: int, seed: int):
        super().__init__()
        g = torch.Generator().manual_seed(9500 + seed)
        self.heads = heads
        self.wq = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)
        self.wk = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)
        self.wv = torch.nn.Parameter(torch.randn(dim, dim, generator=g) / dim ** 0.5)

    def forward(self, x, cls_tokens, policy=None):
        u = torch.cat((cls_tokens, x), dim=1)
        B, T, C = u.shape
        H, hd = self.heads, C // self.heads
        q = (cls_tokens @ self.wq).reshape(B, 1, H, hd).transpose(1, 2)
        k = (u @ self.wk).reshape(B, T, H, hd).transpose(1, 2)
        v = (u @ self.wv).reshape(B, T, H, hd).transpose(1, 2)
        logits = (q @ k.transpose(-2, -1)) * (3.0 * hd ** -0.5)                    # (B,H,1,T)
        e = torch.exp(logits - logits.max(dim=-1, keepdim=True)[0])
        if policy is not None:
            e = e * policy.reshape(B, 1, 1, T)
        attn = e / e.sum(dim=-1, keepdim=True)
        return cls_tokens + 0.5 * (attn @ v).transpose(1, 2).reshape(B, 1, C), attn


class _PatchOnly(torch.nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.blk = blk

    def forward(self, x):
        B, T, _ = x.shape
        return self.blk(x, torch.ones(B, T, 1, device=x.device))


class FakeCait(torch.nn.Module):
    """``blocks`` / ``blocks_token_only`` / ``norm`` / ``layer_nums``: what MyCait.forward_feature_mask_train_direct uses."""

    def __init__(self, dim: int, heads: int, depth: int, depth_token_only: int):
        super().__init__()
        self.blocks = torch.nn.ModuleList([_PatchOnly(FakeAttnBlock(dim, heads, i)) for i in range(depth)])
        self.blocks_token_only = torch.nn.ModuleList([FakeClassAttnBlock(dim, heads, i) for i in range(depth_token_only)])
        self.norm = torch.nn.LayerNorm(dim)
        self.layer_nums = [depth, depth_token_only]
