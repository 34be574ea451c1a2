"""Golden fixtures for the class-row maps from the reference script's OWN lines.

    python tests/golden/make_classmap_golden.py          (build container only: needs /root/reference)

oracle/ref_script_lines.py executes eval_interpretability.py:196-202 and :213-225 as written on seeded inputs; the
result (B, 10, side, side) is stored with fingerprints of the inputs.  The reference hard-codes 10 prototypes per class
(:198), so every case has m = 10.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_script_lines as S  # noqa: E402
from protopformer_b200 import synth  # noqa: E402

# name: (B, C, K, N, seed)
CASES = {"classmap_small": (6, 7, 25, 49, 1), "classmap_cub_b4": (4, 200, 81, 196, 2), "classmap_cars_b3": (3, 196, 121, 196, 3)}


def make_inputs(B, C, K, N, seed):
    g = torch.Generator().manual_seed(7000 + seed)
    side = int(round(K ** 0.5))
    proto_acts = torch.rand(B, C * 10, side, side, generator=g)
    scores = torch.stack([(torch.randperm(N, generator=g) + 1).float() for _ in range(B)]) / float(N * (N + 1) // 2)
    targets = torch.randint(0, C, (B,), generator=g)
    return proto_acts, scores, targets


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (B, C, K, N, seed) in CASES.items():
        proto_acts, scores, targets = make_inputs(B, C, K, N, seed)
        maps = S.class_maps_by_reference_lines(proto_acts, scores, targets, K, 10)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), maps=maps.astype(np.float32),
                            chk_acts=synth.checksum(proto_acts), chk_scores=synth.checksum(scores),
                            targets=targets.numpy())
        print(name, maps.shape, float(maps.sum()))


if __name__ == "__main__":
    main()
