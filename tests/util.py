"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from oracle import synth

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
}


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
