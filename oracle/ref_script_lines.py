"""Runs the class-map lines of the reference's interpretability script AS WRITTEN (test infrastructure, build container
only: reads /root/reference).

`eval_interpretability.py` cannot be imported (argparse, dataset and checkpoint loading at import time), but the two
blocks that define the class-row maps are plain tensor code:

  :196-202   per batch: gather the `proto_per_class` prototypes of each image's label from `proto_acts` (B,P,h,w)
  :213-225   scatter the reserved tokens' activations back onto the grid of all tokens (zeros elsewhere)

This module cuts exactly those source lines out of the file, strips the `.cuda()` calls (no GPU in the build
container; they are device moves, not arithmetic), and executes them in a namespace holding only the variables the
script has in scope at that point.  Nothing is restated: what runs is the reference's text.
"""
from __future__ import annotations

import os
import textwrap
import types

import numpy as np
import torch

SCRIPT = "/root/reference/eval_interpretability.py"
GATHER_LINES = (196, 202)      # fea_size ... proto_acts = torch.gather(...)
SCATTER_LINES = (213, 225)     # if args.reserve_token_nums[0] != 196: ... all_proto_acts = replace_proto_acts


def available() -> bool:
    return os.path.exists(SCRIPT)


def _lines(lo: int, hi: int) -> str:
    src = open(SCRIPT).read().splitlines()[lo - 1:hi]
    return textwrap.dedent("\n".join(src)).replace(".cuda()", "")


def class_maps_by_reference_lines(proto_acts: torch.Tensor, token_attn: torch.Tensor, targets: torch.Tensor,
                                  token_reserve_num: int, num_prototypes_per_class: int) -> np.ndarray:
    """proto_acts (B,P,h,w) as `push_forward` returns it, token_attn (B,N), targets (B,) -> (B, 10, side, side)."""
    gather_src, scatter_src = _lines(*GATHER_LINES), _lines(*SCATTER_LINES)
    assert "torch.gather(proto_acts, 1, proto_indices)" in gather_src and "scatter_(2, reserve_token_indices" in scatter_src, \
        "the reference script's lines moved: update GATHER_LINES / SCATTER_LINES"
    ns = dict(torch=torch, np=np, proto_acts=proto_acts.clone(), targets=targets.clone())
    exec(compile(gather_src, SCRIPT + f":{GATHER_LINES[0]}", "exec"), ns)
    # the script concatenates the per-batch results and goes through numpy (:209-210)
    ns2 = dict(torch=torch, np=np, all_proto_acts=ns["proto_acts"].cpu().detach().numpy(),
               all_token_attn=token_attn.clone(), token_reserve_num=token_reserve_num,
               num_prototypes_per_class=num_prototypes_per_class,
               args=types.SimpleNamespace(reserve_token_nums=[token_reserve_num]))
    exec(compile(scatter_src, SCRIPT + f":{SCATTER_LINES[0]}", "exec"), ns2)
    return np.asarray(ns2["all_proto_acts"])
