"""Golden fixtures for the attention-rollout row from the UNMODIFIED reference.

    python tests/golden/make_rollout_golden.py          (build container only: needs /root/reference)

For each case the reference's own ``MyVisionTransformer.attn_rollout`` (tools/deit_models_attn.py:99-124) is run on
CPU fp32 on seeded synthetic attention maps (oracle/rollout_oracle.synth_attention) and
``attn_rollout[:, 0, 1:]`` (the consumer's view, :226) is stored, plus the top-K selection the head derives from it
(protopformer.py:157-158) and fingerprints of the inputs.  Cases with a tie at the discard threshold are refused
(ATen's topk leaves the choice among equal values unspecified).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness, rollout_oracle as R  # noqa: E402
from protopformer_b200 import synth  # noqa: E402

# name: (L, B, H, T, seed, head_fusion, K)
CASES = {
    "rollout_tiny": (3, 2, 2, 12, 1, "mean", 3),                 # sparse scores: only the 3 largest are tie-free
    "rollout_small_h6": (4, 3, 6, 50, 2, "mean", 25),
    "rollout_max_fusion": (3, 2, 4, 30, 3, "max", 9),
    "rollout_deit_tiny_b2": (11, 2, 3, 197, 4, "mean", 81),       # DeiT-Ti: 3 heads, reserve layer 11 (train_cub.sh)
}


# name: (n_patch, n_cls, B, H, T, seed, K)   -- tools/cait_models_attn.py:223-261 (cait_xxs24: 24 patch layers, 4 heads)
CAIT_CASES = {
    "rollout_cait_tiny": (3, 1, 2, 2, 16, 1, 2),
    "rollout_cait_xxs24_b2": (24, 2, 2, 4, 196, 2, 121),
}


def cait():
    import tools.cait_models_attn as cm
    for name, (n_patch, n_cls, B, H, T, seed, K) in CAIT_CASES.items():
        attn = R.synth_cait_attention(n_patch, n_cls, B, H, T, seed)
        assert R.threshold_tie_free(attn, 0.9, "mean"), f"{name}: tie at the discard threshold"
        _, cls_result = cm.MyCait.attn_rollout_cait(None, [a.clone() for a in attn], discard_ratio=0.9,
                                                    head_fusion="mean", layer_nums=[n_patch, n_cls])
        scores = cls_result[:, 0].contiguous()                      # cls_attn_ma[:, 0], cait_models_attn.py:330
        idx = torch.topk(scores, k=K, dim=-1)[1].sort(dim=-1)[0]
        srt = scores.sort(dim=-1, descending=True)[0]
        out = dict(scores=scores.numpy(), idx=idx.numpy().astype(np.int32),
                   sel_gap=np.float32(((srt[:, K - 1] - srt[:, K]) / srt[:, K - 1].clamp_min(1e-30)).min().item()),
                   chk=np.array([synth.checksum(a) for a in attn]))
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
        print(name, "scores", tuple(scores.shape), "relative selection gap %.2e" % float(out["sel_gap"]))


def main():
    ref_harness.import_reference()
    import tools.deit_models_attn as dm          # the reference's module (timm stubbed by ref_harness)
    torch.set_num_threads(1)
    for name, (L, B, H, T, seed, fusion, K) in CASES.items():
        attn = R.synth_attention(L, B, H, T, seed)
        assert R.threshold_tie_free(attn, 0.9, fusion), f"{name}: tie at the discard threshold, pick another seed"
        full = dm.MyVisionTransformer.attn_rollout(None, [a.clone() for a in attn], discard_ratio=0.9,
                                                   head_fusion=fusion)
        scores = full[:, 0, 1:].contiguous()
        idx = torch.topk(scores, k=K, dim=-1)[1].sort(dim=-1)[0]
        srt = scores.sort(dim=-1, descending=True)[0]
        out = dict(scores=scores.numpy(), idx=idx.numpy().astype(np.int32),
                   sel_gap=np.float32(((srt[:, K - 1] - srt[:, K]) / srt[:, K - 1].clamp_min(1e-30)).min().item()),
                   row_sum=full[:, 0].sum(-1).numpy(),
                   chk=np.array([synth.checksum(a) for a in attn]))
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
        print(name, "scores", tuple(scores.shape), "relative selection gap %.2e" % float(out["sel_gap"]))
    cait()


if __name__ == "__main__":
    main()
